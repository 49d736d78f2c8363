#!/bin/bash
MOS_FUSION_PROFILE=1 timeout 600 python tools/config_bench.py fusion 2>&1 | tail -3
