"""B200 mirror of the reference training loop (train_edlora.py:105-158) for the UNet LoRA group: one process per GPU,
the batch sharded across ranks, ONE all-reduce per optimiser step on the flat LoRA gradient (+ loss), fused AdamW,
linear learning-rate decay to zero (diffusers get_scheduler('linear', warmup 0), train_edlora.py:85-90).

The data pipeline (LoraDataset, VAE, CLIP) is out of scope / next: `train` consumes an iterator of
already-encoded batches  dict(latents, encoder_hidden_states, masks, img_masks[, text_input_ids]).
"""
import torch
import torch.distributed as dist

from mos_b200.dp import allreduce_flat, optimizer_step


def total_iterations(dataset_len, batch_size_per_gpu, world_size, gradient_accumulation_steps=1):
    """train_edlora.py:73-75 (a float, compared with `global_step < total_iter`)."""
    return dataset_len / (batch_size_per_gpu * world_size * gradient_accumulation_steps)


def linear_lr(base_lr, step, num_training_steps):
    """LambdaLR of diffusers' linear schedule with 0 warm-up: lr_k = base * max(0, (T - k) / T)."""
    return base_lr * max(0.0, float(num_training_steps - step) / float(max(1.0, num_training_steps)))


def train(trainer, batches, *, dataset_len, batch_size_per_gpu, gradient_accumulation_steps=1, print_freq=0,
          log=print):
    """Runs the loop of train_edlora.py:105-158; returns the list of per-step mean losses (rank-averaged)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    total_iter = total_iterations(dataset_len, batch_size_per_gpu, world, gradient_accumulation_steps)
    sched_steps = total_iter * gradient_accumulation_steps
    eng = trainer.engine
    state = eng.state
    base_lr = trainer.unet_lr
    it = iter(batches)
    global_step, micro, sched_k = 0, 0, 0
    losses = []
    while global_step < total_iter:
        batch = next(it)
        masks = batch['masks'] if 'masks' in batch else batch['img_masks']
        loss = trainer(batch['latents'], batch['encoder_hidden_states'], masks, batch['img_masks'],
                       text_input_ids=batch.get('text_input_ids'), accumulate=micro > 0)
        micro += 1
        # accelerate steps the optimiser (and the scheduler) on every micro-batch call but only applies it when
        # gradients are synchronised; the schedule therefore advances once per micro-step (train_edlora.py:128-130)
        if micro == gradient_accumulation_steps:
            state.lrs = (state.lrs[0], state.lrs[1], linear_lr(base_lr, sched_k, sched_steps))
            grad_scale, mean_loss, _ = allreduce_flat(state, loss_value=float(loss))
            optimizer_step(state, grad_scale / gradient_accumulation_steps)
            eng.refresh_lora()
            micro = 0
            global_step += 1
            losses.append(mean_loss)
            if print_freq and global_step % print_freq == 0:
                log(f'iter {global_step}: loss {mean_loss:.5f} lr {state.lrs[2]:.3e}')
        sched_k += 1
    return losses
