"""Worker of tests/test_multigpu.py (launched by torchrun, one rank per GPU). Every sharded path of SURVEY.md §8e is compared with the
CPU ORACLE evaluated on the GATHERED (global) batch — the reference's own semantics, where every rank holds the full batch after
gather_tensor (antmmf/utils/distributed_utils.py:145-189) and evaluates the full loss:

  clip      CNCLIP.contrastive_loss (symmetric InfoNCE, cn_model.py:221-223): loss and every parameter gradient after the DDP mean
  mil       get_mil_nce_loss, n_clips = 1 (univl_video_ret.py:146-197): loss and feature gradients
  mil_n2    the same with two clips per video (forward_stage1, :357-387)
  moco      MoCo key all-gather + enqueue (moco_utils.py:84-108): queue contents / pointer bit-exact, queue NCE loss vs the oracle

Prints one `MGPU <case> ...` line per case and `MGPU ALL OK` when every case is within its bound; exits non-zero otherwise."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

BF = torch.bfloat16


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-12))


def case_clip(rank, world, restated):
    """Sharded model step vs the fp32 oracle on the global batch (bf16-rounded weights / inputs)."""
    from b200mm.modules import CNCLIP

    cfg = dict(embed_dim=64, image_resolution=32, vision_layers=2, vision_width=128, vision_patch_size=8, vocab_size=300,
               text_attention_probs_dropout_prob=0.0, text_hidden_act="gelu", text_hidden_dropout_prob=0.0, text_hidden_size=128,
               text_initializer_range=0.02, text_intermediate_size=256, text_max_position_embeddings=32, text_num_attention_heads=2,
               text_num_hidden_layers=2, text_type_vocab_size=2, vision_head_width=64)
    torch.manual_seed(0)
    model = CNCLIP(**cfg)
    sd = model.state_dict()
    model = model.cuda().to(BF).train()
    Bl = 12
    g = torch.Generator().manual_seed(5)
    image = torch.randn(Bl * world, 3, 32, 32, generator=g)
    text = torch.randint(1, 300, (Bl * world, 16), generator=g)
    text[:, 0] = 101
    text[::3, 9:] = 0
    sl = slice(rank * Bl, (rank + 1) * Bl)
    loss_local = model.contrastive_loss(image[sl].cuda(), text[sl].cuda())
    loss_local.backward()
    lt = loss_local.detach().float().clone()
    dist.all_reduce(lt)
    lt /= world
    grads = {}
    for n, p in model.named_parameters():
        gr = p.grad.float().clone()
        dist.all_reduce(gr)
        grads[n] = gr / world
    # oracle on the global batch
    sd16 = {k: (v.to(BF).float().clone().requires_grad_(True) if torch.is_floating_point(v) else v) for k, v in sd.items()}
    _, _, logits, _ = restated.cnclip_forward(sd16, image.to(BF).float(), text, 2, 2)
    ref = restated.symmetric_info_nce(logits)
    ref.backward()
    # calibrator (as in tests/test_model_gpu.py): the same oracle arithmetic run in bf16 by torch eager on this GPU, global batch
    sdb = {k: (v.detach().to(BF).cuda().requires_grad_(True) if torch.is_floating_point(v) else v.cuda()) for k, v in sd.items()}
    _, _, e_logits, _ = restated.cnclip_forward(sdb, image.cuda().to(BF), text.cuda(), 2, 2)
    restated.symmetric_info_nce(e_logits.float()).backward()
    ok = abs(float(lt) - float(ref)) < 2e-3 * max(1.0, abs(float(ref)))
    worst, worst_n, worst_ratio = 0.0, "", 0.0
    for n in grads:
        r = sd16[n].grad
        if r is None or float(r.abs().max()) < 1e-6 or n == "logit_scale":
            continue
        e, e_eager = rel_l2(grads[n], r), rel_l2(sdb[n].grad, r)
        if e > worst:
            worst, worst_n = e, n
        worst_ratio = max(worst_ratio, e / max(e_eager, 2e-2))
        ok = ok and e < max(4e-2, 2.0 * e_eager)
    return ok, (f"loss_sharded={float(lt):.5f} oracle={float(ref):.5f} worst_grad_rel_l2={worst:.3e} ({worst_n}) "
                f"worst ours/max(eager_bf16, 2e-2)={worst_ratio:.2f}")


def case_mil(rank, world, restated, n_clips):
    from b200mm.contrastive import mil_nce_loss

    Bl, E = 10, 64  # global batch 10 * world: not a multiple of 8 for world = 2 (exercises the zero-row padding of the gathered operand)
    g = torch.Generator().manual_seed(11 + n_clips)
    t_all = torch.nn.functional.normalize(torch.randn(Bl * world, E, generator=g), dim=-1).to(BF)
    v_all = torch.nn.functional.normalize(torch.randn(Bl * world * n_clips, E, generator=g), dim=-1).to(BF)
    t = t_all[rank * Bl:(rank + 1) * Bl].cuda().requires_grad_()
    v = v_all[rank * Bl * n_clips:(rank + 1) * Bl * n_clips].cuda().requires_grad_()
    loss = mil_nce_loss(v, t, n_clips=n_clips)
    loss.backward()
    lt = loss.detach().float().clone()
    dist.all_reduce(lt)
    lt /= world
    tf, vf = t_all.float().requires_grad_(), v_all.float().requires_grad_()
    sim = restated.l1_simi_matrix(tf, vf, n_clips)
    ref = restated.mil_nce_n1(sim.view(tf.shape[0], -1)) if n_clips == 1 else restated.mil_nce_clips(sim)
    ref.backward()
    # every rank back-propagates its own W * share: the local-row gradient is W x the gradient of the global loss
    et = rel_l2(t.grad / world, tf.grad[rank * Bl:(rank + 1) * Bl])
    ev = rel_l2(v.grad / world, vf.grad[rank * Bl * n_clips:(rank + 1) * Bl * n_clips])
    ok = abs(float(lt) - float(ref)) < 1e-4 * max(1.0, abs(float(ref))) and et < 1e-2 and ev < 1e-2
    return ok, f"n_clips={n_clips} loss_sharded={float(lt):.6f} oracle={float(ref):.6f} dtext_rel_l2={et:.3e} dvideo_rel_l2={ev:.3e}"


def case_moco(rank, world, restated):
    from b200mm.moco import B200MocoUtils, moco_nce

    E, K, Bl = 64, 256, 12
    enc = torch.nn.Linear(E, E)
    torch.manual_seed(3)
    mo = B200MocoUtils({"hidden_size": E, "K": K, "M": 0.999, "T": 0.05}, img_encoder=enc, txt_encoder=torch.nn.Linear(E, E))
    mo.img_K = K
    mo.img_queue = torch.nn.functional.normalize(torch.randn(E, K), dim=0)
    mo = mo.cuda()
    g = torch.Generator().manual_seed(21)
    keys_v = torch.nn.functional.normalize(torch.randn(Bl * world, E, generator=g), dim=-1).to(BF)
    keys_t = torch.nn.functional.normalize(torch.randn(Bl * world, E, generator=g), dim=-1).to(BF)
    q_all = torch.nn.functional.normalize(torch.randn(Bl * world, E, generator=g), dim=-1).to(BF)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    before_t = mo.txt_queue.clone()
    mo.dequeue_and_enqueue(keys_v[sl].cuda(), keys_t[sl].cuda())
    want_t = before_t.clone()
    want_t[:, :Bl * world] = keys_t.t().float().cuda()
    ok = torch.equal(mo.txt_queue, want_t) and int(mo.txt_queue_ptr) == (Bl * world) % K and int(mo.img_queue_ptr) == (Bl * world) % K
    ok = ok and torch.equal(mo.img_queue[:, :Bl * world], keys_v.t().float().cuda())
    # queue NCE on the updated queue: local queries / positives against the (replicated) queue, vs the oracle on this rank's rows
    q = q_all[sl].cuda().requires_grad_()
    loss = moco_nce(q, keys_t[sl].cuda(), mo._queue_bf16("txt_queue"), 0.05)
    loss.backward()
    qf = q_all[sl].float().requires_grad_()
    queue16 = mo.txt_queue.to(BF).float().cpu()
    pos = (qf * keys_t[sl].float()).sum(-1, keepdim=True)
    ref = restated.moco_nce(pos, qf @ queue16, 0.05)
    ref.backward()
    e = rel_l2(q.grad, qf.grad)
    ok = ok and abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref))) and e < 1e-2
    return ok, f"queue/ptr exact={ok} loss={float(loss):.6f} oracle={float(ref):.6f} dq_rel_l2={e:.3e}"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import b200mm  # noqa: F401
    from oracle import restated

    cases = [("clip", lambda: case_clip(rank, world, restated)), ("mil", lambda: case_mil(rank, world, restated, 1)),
             ("mil_n2", lambda: case_mil(rank, world, restated, 2)), ("moco", lambda: case_moco(rank, world, restated))]
    bad = 0
    for name, fn in cases:
        ok, msg = fn()
        flag = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flag)
        bad += int(flag)
        if rank == 0:
            print(f"MGPU {name} world={world} {msg}", "OK" if int(flag) == 0 else f"FAIL (on {int(flag)} ranks)", flush=True)
    if rank == 0 and bad == 0:
        print("MGPU ALL OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
