"""Extract the per-kernel evidence of an `ncu --set full` report into small committed files:
    python tools/ncu_summary.py gpurun_out/r2_kernels.ncu-rep profiles/r2_kernels
-> profiles/r2_kernels_summary.csv (one row per captured launch: duration, DRAM bytes, tensor-pipe / XU / L2 / DRAM
utilisation, registers, achieved occupancy) and profiles/r2_kernels_summary.json (the same, read by bench.py for
roofline.traffic)."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
METRICS = {
    'gpu__time_duration.sum': 'duration_us',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct',
    'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active': 'tensor_inst_pct',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'xu_pipe_pct',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'lts__t_bytes.sum': 'l2_bytes',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_pct',
    'l1tex__m_xbar2l1tex_read_bytes.sum': 'l2_to_sm_read_bytes',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occupancy_pct',
    'launch__registers_per_thread': 'regs',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
}
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(v):
    try:
        return float(v.replace(',', ''))
    except Exception:
        return None


def scale(name, v):
    u = units[col[name]]
    if v is None:
        return None
    if name == 'gpu__time_duration.sum':
        return v / 1e3 if u in ('nsecond', 'ns') else (v * 1e3 if u in ('msecond', 'ms') else v)
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u)
    return v * mult if mult else v


kernels = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    k = {'id': r[col['ID']], 'kernel': r[col['Kernel Name']][:90], 'grid': r[col.get('Grid Size', col['ID'])],
         'block': r[col.get('Block Size', col['ID'])]}
    for m, short in METRICS.items():
        k[short] = scale(m, num(r[col[m]])) if m in col else None
    if k.get('dram_read') is not None and k.get('dram_write') is not None:
        k['dram_bytes'] = k['dram_read'] + k['dram_write']
    kernels.append(k)
with open(out + '_summary.json', 'w') as f:
    json.dump({'source': rep, 'kernels': kernels}, f, indent=1)
with open(out + '_summary.csv', 'w') as f:
    keys = list(kernels[0].keys()) if kernels else []
    w = csv.DictWriter(f, fieldnames=keys)
    w.writeheader()
    for k in kernels:
        w.writerow({kk: (f'{v:.4g}' if isinstance(v, float) else v) for kk, v in k.items()})
print(f'{len(kernels)} launches -> {out}_summary.csv / .json')
