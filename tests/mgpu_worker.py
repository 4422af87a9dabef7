"""Worker of tests/test_multigpu.py (launched by torchrun, one rank per GPU): sharded global-batch contrastive training step
vs the same step on the full batch in one process."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import b200mm
    from b200mm.modules import CNCLIP

    cfg = dict(embed_dim=64, image_resolution=32, vision_layers=2, vision_width=128, vision_patch_size=8, vocab_size=300,
               text_attention_probs_dropout_prob=0.0, text_hidden_act="gelu", text_hidden_dropout_prob=0.0, text_hidden_size=128,
               text_initializer_range=0.02, text_intermediate_size=256, text_max_position_embeddings=32, text_num_attention_heads=2,
               text_num_hidden_layers=2, text_type_vocab_size=2, vision_head_width=64)
    torch.manual_seed(0)
    model = CNCLIP(**cfg).cuda().to(torch.bfloat16).train()
    Bl = 12  # per-rank batch; global batch is not a multiple of 8*world on purpose? 12*2 = 24 -> multiple of 8 ; use mil (pads) too
    g = torch.Generator().manual_seed(5)
    image = torch.randn(Bl * world, 3, 32, 32, generator=g).cuda()
    text = torch.randint(1, 300, (Bl * world, 16), generator=g).cuda()
    text[:, 0] = 101
    text[::3, 9:] = 0
    sl = slice(rank * Bl, (rank + 1) * Bl)

    # --- sharded step (what DDP does: mean of per-rank losses, mean of per-rank gradients)
    loss_local = model.contrastive_loss(image[sl], text[sl])
    loss_local.backward()
    grads = {n: p.grad.float().clone() for n, p in model.named_parameters()}
    for p in model.parameters():
        p.grad = None
    lt = loss_local.detach().float().clone()
    dist.all_reduce(lt)
    lt /= world
    for n in grads:
        dist.all_reduce(grads[n])
        grads[n] /= world

    # --- the same global batch in one process (no process group -> world 1 semantics)
    from b200mm.contrastive import clip_contrastive_loss

    class _Solo:  # a group-like sentinel is not needed: compute with the collective-free path by hiding the process group
        pass

    img, txt = model.encode_normalized(image, text)
    import b200mm.contrastive as C
    real = C._world
    C._world = lambda group=None: (0, 1)
    try:
        loss_full = clip_contrastive_loss(img, txt, model.logit_scale)
        loss_full.backward()
    finally:
        C._world = real
    ok = True
    msg = []
    if abs(float(lt) - float(loss_full)) > 2e-3 * max(1.0, abs(float(loss_full))):
        ok = False
        msg.append(f"loss sharded {float(lt)} vs full {float(loss_full)}")
    worst = 0.0
    for n, p in model.named_parameters():
        ref = p.grad.float()
        sc = float(ref.abs().max())
        if sc < 1e-6:
            continue
        e = float((grads[n] - ref).norm() / ref.norm().clamp_min(1e-12))
        worst = max(worst, e)
        if e > 4e-2:
            ok = False
            msg.append(f"{n}: rel-l2 {e:.3e}")
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print(f"MGPU loss_sharded={float(lt):.5f} loss_full={float(loss_full):.5f} worst_grad_rel_l2={worst:.3e}", "OK" if int(flag) == 0 else "FAIL " + "; ".join(msg[:5]), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 0 else 1)


if __name__ == "__main__":
    main()
