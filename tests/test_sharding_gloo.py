"""CPU, world_size 2 over gloo: the cutout-sharded algorithm of the multi-GPU path, executed with the
oracle standing in for the kernels.  Each rank owns cutouts [r*cutn/R, (r+1)*cutn/R), the drawer is replicated, the
global min/max of the range normalise is exchanged with one allreduce(min) over {min, -max}, the d/dmin, d/dmax sums
and the IMAGE gradient are summed with allreduce -- exactly the collectives the engine issues (engine.cu:
forward_cutouts, backward_all).  The resulting z.grad must equal the single-process gradient.

The exchange point is the image gradient, not z.grad: ClampWithGrad's backward (vqgan.py:76-79) masks by the SIGN of
the incoming gradient, so the drawer backward is not linear in it and sum_r bwd(g_r) != bwd(sum_r g_r)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_path as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup():
    torch.manual_seed(0)
    vq = R.init_vqgan_weights(R.VQModel(n_embed=64, embed_dim=32, ch=32, ch_mult=(1, 2), num_res_blocks=1,
                                        attn_resolutions=(4,), resolution=8, z_channels=32), 4)
    clip = R.init_clip_weights(R.ClipVisual(32, 8, 64, 2, 1, 16), 3)
    g = torch.Generator().manual_seed(1)
    z = vq.quantize.embedding.weight[torch.randint(64, (16,), generator=g)].T.reshape(1, 32, 4, 4).clone()
    cutn, cs = 6, 32
    T = torch.eye(3).repeat(cutn, 1, 1)
    T[:, 0, 0] = torch.tensor([1.3, 1.1, 1.5, 0.9, 0.95, 0.85])
    T[:, 1, 1] = torch.tensor([1.2, 1.4, 1.1, 0.9, 0.8, 0.9])
    T[:, 0, 2] = torch.tensor([-3.0, 1.5, -6.0, 1.0, 2.0, 0.5])
    facs = torch.rand(cutn, generator=g) * 0.1
    noise = torch.randn(cutn, 3, cs, cs, generator=g)
    prompts = [(torch.randn(1, 16, generator=g), 1.0, float("-inf")), (torch.randn(1, 16, generator=g), -0.4, float("-inf"))]
    return vq, clip, z, cutn, cs, T, facs, noise, prompts


def _extras():
    """ColorJitter rows (per GLOBAL cutout index) and two image prompts for the second test."""
    from pixray_b200 import cutouts
    J = torch.from_numpy(cutouts.sample_color_jitter(6, 5, p=1.0))
    J[1, 0] = 0
    g = torch.Generator().manual_seed(9)
    targets = [(torch.rand(1, 3, 8, 8, generator=g), 0.6), (torch.rand(1, 3, 8, 8, generator=g), -0.3)]
    return J, targets


def _local_cutouts(pooled, T, lo, n_local, zoom, cs, facs, noise, J=None):
    parts = []
    for n in range(lo, lo + n_local):
        if n < zoom:
            parts.append(R.warp_perspective(pooled, T[n:n + 1], (cs, cs), padding_mode="reflection"))
        else:
            parts.append(R.warp_perspective(pooled, T[n:n + 1], (cs, cs), padding_mode="fill", fill_value=[0.4] * 3))
    batch = torch.cat(parts)
    if J is not None:
        batch = R.color_jitter(batch, J[lo:lo + n_local])
    return batch + facs[lo:lo + n_local].reshape(-1, 1, 1, 1) * noise[lo:lo + n_local]


def _sharded_grad(rank, world, port, out, extras=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    vq, clip, z, cutn, cs, T, facs, noise, prompts = _setup()
    n_local = cutn // world
    lo = rank * n_local
    zz = z.clone().requires_grad_(True)
    img_full = R.vqgan_synth(vq, zz)                                      # replicated drawer
    img = img_full.detach().requires_grad_(True)
    pooled = R.pool_avg_max(img, cs)
    zoom = int(0.6 * cutn)                                                # group split by GLOBAL index
    J, targets = _extras() if extras else (None, [])
    mean = torch.tensor(R.CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(R.CLIP_STD).view(1, 3, 1, 1)
    # image prompts first (engine.cu: encode_image_prompts): every rank cuts and encodes ITS slice of each target with the
    # iteration's transforms (no ColorJitter on that path), the range normalise of the target batch is global too, and the
    # [cutn, D] rows are completed with one allreduce(sum) over a buffer that is zero outside the rank's own rows
    image_prompts = []
    for (timg, w) in targets:
        with torch.no_grad():
            tb = _local_cutouts(R.pool_avg_max(timg, cs), T, lo, n_local, zoom, cs, facs, noise)
            x = torch.stack([tb.min(), -tb.max()])
            dist.all_reduce(x, op=dist.ReduceOp.MIN)
            te = clip.encode_image(((tb - x[0]) / (-x[1] - x[0]) - mean) / std)
            te = te / te.norm(dim=-1, keepdim=True)
            rows = torch.zeros(cutn, te.shape[1])
            rows[lo:lo + n_local] = te
            dist.all_reduce(rows, op=dist.ReduceOp.SUM)
        image_prompts.append((rows, w, float("-inf")))
    batch = _local_cutouts(pooled, T, lo, n_local, zoom, cs, facs, noise, J)
    # global range: one allreduce(min) over {min, -max}; autograd ownership of the extreme element stays local
    lmin, lmax = batch.min(), batch.max()
    x = torch.stack([lmin.detach(), -lmax.detach()])
    dist.all_reduce(x, op=dist.ReduceOp.MIN)
    gmin, gmax = x[0], -x[1]
    # The range scalars are shared by every rank's cutouts: treat them as leaves, exchange their gradients
    # (allreduce(sum) -- the engine's `sums` exchange), then the rank owning the extreme element routes the total
    # into its element (what range_unpack's -1 index + cutout_backward's dMin / dR terms implement).
    mn = gmin.clone().requires_grad_(True)
    mx = gmax.clone().requires_grad_(True)
    a = batch - mn
    y = a / (mx - mn)
    e = clip.encode_image((y - mean) / std)
    e = e / e.norm(dim=-1, keepdim=True)
    loss = 0
    for (embed, w, stop) in list(prompts) + image_prompts:
        # Prompt.forward's mean runs over ALL cutn cutouts: each rank contributes sum / cutn_global
        loss = loss + R.prompt_loss(e, embed, w, stop) * (n_local / cutn)
    loss.backward(retain_graph=True)
    d = torch.stack([mn.grad, mx.grad])
    dist.all_reduce(d, op=dist.ReduceOp.SUM)
    if lmin.detach() == gmin:
        lmin.backward(d[0], retain_graph=True)
    if lmax.detach() == gmax:
        lmax.backward(d[1])
    g_img = img.grad.clone()
    dist.all_reduce(g_img, op=dist.ReduceOp.SUM)                           # the path's one exchange step
    img_full.backward(g_img)                                               # replicated drawer backward on the full grad
    g = zz.grad.clone()
    l = loss.detach().clone()
    dist.all_reduce(l, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save(dict(grad=g, loss=l), out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_cutout_sharding_matches_single_process(tmp_path):
    vq, clip, z, cutn, cs, T, facs, noise, prompts = _setup()
    ref = R.iterate(lambda zz: R.vqgan_synth(vq, zz), z, [clip], [prompts], T, cs, "reflection", 0.4, facs, noise)
    out = str(tmp_path / "sharded.pt")
    port = _free_port()
    mp.spawn(_sharded_grad, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    err = (got["grad"] - ref["z_grad"]).abs().max().item()
    mag = ref["z_grad"].abs().max().item()
    assert mag > 0 and err <= 2e-5 * max(1.0, mag), (err, mag)
    assert abs(float(got["loss"]) - float(sum(ref["losses"]))) < 1e-5


@pytest.mark.timeout(300)
def test_sharding_with_image_prompts_and_color_jitter(tmp_path):
    """The two additions to the exchange: ColorJitter rows are per global cutout index (nothing to exchange), image
    prompts need all cutn target rows on every rank (one allreduce per target and perceptor)."""
    vq, clip, z, cutn, cs, T, facs, noise, prompts = _setup()
    J, targets = _extras()
    ref = R.iterate(lambda zz: R.vqgan_synth(vq, zz), z, [clip], [prompts], T, cs, "reflection", 0.4, facs, noise,
                    jitter=J, image_prompts=targets)
    out = str(tmp_path / "sharded.pt")
    mp.spawn(_sharded_grad, args=(2, _free_port(), out, True), nprocs=2, join=True)
    got = torch.load(out)
    err = (got["grad"] - ref["z_grad"]).abs().max().item()
    mag = ref["z_grad"].abs().max().item()
    assert mag > 0 and err <= 2e-5 * max(1.0, mag), (err, mag)
    assert abs(float(got["loss"]) - float(sum(ref["losses"]))) < 1e-5
