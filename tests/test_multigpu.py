"""GPU, 2 ranks over NCCL: the cutout-sharded mode.  Each rank owns half of the cutouts; the engine exchanges the
global min/max, the range-gradient sums and the image gradient; z.grad on every rank must equal the single-GPU oracle
gradient, and be bit-identical across ranks (replicated drawer backward on identical bits)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, with_aux=0):
    import torch.distributed as dist
    from test_pipeline_gpu import SMALL_CLIP, SMALL_VQ, plant_extremes, random_transforms
    from oracle import ref_path as R
    from pixray_b200 import engine as E
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cutn, cs, seed = 8, 224, 21
    vq = R.init_vqgan_weights(R.VQModel(n_embed=1024, embed_dim=128, ch=128, ch_mult=(1, 2), num_res_blocks=1,
                                        attn_resolutions=(16,), resolution=32, z_channels=128), seed)
    clip = R.init_clip_weights(R.ClipVisual(224, 32, 128, 2, 2, 64), seed + 1)
    eng = E.B200Engine(drawer=E.DRAWER_VQGAN, image_hw=(32, 32), vqgan=SMALL_VQ, cutn=cutn, clip=[SMALL_CLIP], seed=seed,
                       device=rank, rank=rank, world=world)
    eng.load_module(E.MOD_VQGAN, vq.state_dict())
    eng.load_module(E.MOD_CLIP0, clip.state_dict())
    eng.finalize()
    eng.init_comm()
    g = torch.Generator().manual_seed(seed + 2)
    prompts = [(torch.randn(1, 64, generator=g), 1.0, float("-inf")), (torch.randn(1, 64, generator=g), -0.3, float("-inf"))]
    eng.set_prompts(0, torch.cat([p[0] for p in prompts]).numpy(), [1.0, -0.3], [float("-inf")] * 2)
    idx = torch.randint(1024, (256,), generator=g)
    z = (vq.quantize.embedding.weight[idx].T.reshape(1, 128, 16, 16) + 0.05 * torch.randn(1, 128, 16, 16, generator=g)).contiguous()
    T = random_transforms(cutn, cs, 3)
    facs, noise = plant_extremes(torch.rand(cutn, generator=g) * 0.1, torch.randn(cutn, 3, cs, cs, generator=g))
    aux = []
    if with_aux == 1:
        # auxiliary losses under sharding: saturation needs the GLOBAL colour moments, smoothness the neighbouring ranks'
        # boundary rows (the reference differentiates across the stacked cutouts), the image losses are replicated
        pal = [[0.9, 0.1, 0.1], [0.1, 0.8, 0.2], [0.2, 0.2, 0.9]]
        hw_ = torch.randn(1, 64, generator=torch.Generator().manual_seed(5)) * 0.3
        eng.add_aux_loss(E.LOSS_SATURATION, 0.5, [1.3])
        eng.add_aux_loss(E.LOSS_SMOOTHNESS, 2.0, [0.9, 0, 1])
        eng.add_aux_loss(E.LOSS_PALETTE, 1.0, [0.8] + [c for row in pal for c in row])
        eng.add_aux_loss(E.LOSS_SYMMETRY, 1.5, [0.7])
        eng.add_aux_loss(E.LOSS_AESTHETIC, 0.8, [10.0, 0.7] + hw_.reshape(-1).tolist())
        aux = [(0.5, lambda o, b, e: R.saturation_loss(b, 1.3)), (2.0, lambda o, b, e: R.smoothness_loss(b, 0.9)),
               (1.0, lambda o, b, e: R.palette_loss(b, pal, 0.8)[0]), (1.5, lambda o, b, e: R.symmetry_loss(o, 0.7)),
               (0.8, lambda o, b, e: R.aesthetic_loss(e, hw_, torch.tensor([0.7]), 10.0))]
    jitter, targets = None, []
    if with_aux == 2:
        # image prompts (every rank needs all cutn rows of the target embeddings: one more allreduce) and the ColorJitter
        # rows (per global cutout index) under sharding
        from pixray_b200 import cutouts
        aux = []
        jitter = cutouts.sample_color_jitter(cutn, 31, p=1.0)
        tg = torch.Generator().manual_seed(33)
        targets = [(torch.rand(1, 3, 32, 32, generator=tg), 0.7), (torch.rand(1, 3, 32, 32, generator=tg) * 0.6, -0.5)]
        eng.set_image_prompts(torch.cat([t for t, _ in targets]), [w for _, w in targets])
    zc = z.clone().cuda()
    losses = np.zeros(2 + len(aux) + len(targets), dtype=np.float32)
    eng.iterate(zc, 0.05, 0, params=dict(transforms=T, zoom_padding=0, fill=0.4, noise_facs=facs.numpy(), noise=noise,
                                         color_jitter=jitter), losses_out=losses)
    zg = eng.debug_read("z_grad", z.shape).cpu()
    torch.save(dict(zg=zg, losses=losses.copy(), z=zc.cpu()), os.path.join(out_dir, f"rank{rank}.pt"))
    if rank == 0:
        ref = R.iterate(lambda zz: R.vqgan_synth(vq, zz), z, [clip], [prompts], torch.from_numpy(T), cs, "reflection", 0.4,
                        facs, noise, aux=aux, jitter=None if jitter is None else torch.from_numpy(jitter),
                        image_prompts=targets)
        torch.save(dict(zg=ref["z_grad"], losses=torch.stack([l.reshape(()) for l in ref["losses"]])),
                   os.path.join(out_dir, "ref.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cutout_sharded_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1, ref = (torch.load(tmp_path / n, weights_only=False) for n in ("rank0.pt", "rank1.pt", "ref.pt"))
    assert torch.equal(r0["zg"], r1["zg"]), "replicated drawer backward must be bit-identical across ranks"
    assert torch.equal(r0["z"], r1["z"])
    err = (r0["zg"] - ref["zg"]).abs().max().item()
    mag = ref["zg"].abs().max().item()
    print(f"[parity] 2-rank sharded z.grad: max_abs_err={err:.3e} ref_max={mag:.3e}; losses {r0['losses']} vs {ref['losses']}")
    assert err <= 3e-2 * mag
    assert np.abs(r0["losses"] - ref["losses"].numpy()).max() < 5e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cutout_sharded_two_ranks_with_aux_losses(tmp_path):
    """Same, with auxiliary losses: their values and gradients must not depend on how the cutouts are sharded."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), 1), nprocs=2, join=True)
    r0, r1, ref = (torch.load(tmp_path / n, weights_only=False) for n in ("rank0.pt", "rank1.pt", "ref.pt"))
    assert torch.equal(r0["zg"], r1["zg"]) and torch.equal(r0["z"], r1["z"])
    err = (r0["zg"] - ref["zg"]).abs().max().item()
    mag = ref["zg"].abs().max().item()
    print(f"[parity] 2-rank sharded z.grad with aux losses: max_abs_err={err:.3e} ref_max={mag:.3e}; losses {r0['losses']} vs {ref['losses']}")
    assert err <= 3e-2 * mag
    assert np.abs(r0["losses"] - ref["losses"].numpy()).max() < 5e-3
    assert np.array_equal(r0["losses"], r1["losses"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cutout_sharded_two_ranks_with_image_prompts_and_color_jitter(tmp_path):
    """Image prompts + ColorJitter under sharding: loss vector as on one GPU, ranks bit-identical.  z.grad at the
    whole-chain bound of the jittered path (tests/test_color_jitter.py)."""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), 2), nprocs=2, join=True)
    r0, r1, ref = (torch.load(tmp_path / n, weights_only=False) for n in ("rank0.pt", "rank1.pt", "ref.pt"))
    assert torch.equal(r0["zg"], r1["zg"]) and torch.equal(r0["z"], r1["z"])
    assert np.array_equal(r0["losses"], r1["losses"]) and r0["losses"].size == 4
    err = (r0["zg"] - ref["zg"]).abs().max().item()
    mag = ref["zg"].abs().max().item()
    print(f"[parity] 2-rank sharded z.grad with image prompts + jitter: max_abs_err={err:.3e} ref_max={mag:.3e}; losses {r0['losses']} vs {ref['losses']}")
    assert err <= 4e-2 * mag
    assert np.abs(r0["losses"] - ref["losses"].numpy()).max() < 5e-3
