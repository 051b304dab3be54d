"""train()'s control flow on the device (pixray.py:1090-1109 checkdrop, 1464-1512 scheduled drops / auto-stop / rebuild,
1464-1482 batches): the managed optimiser step against the reference's host-side policy replayed on the losses the
engine itself reported, and gradient accumulation over several passes against the sum of single passes."""
import numpy as np
import pytest
import torch

from pixray_b200 import engine as E

pytestmark = pytest.mark.gpu


def host_policy(history, base_lr, delay, max_drops, auto_stop, drops):
    """The reference's checkdrop / train() bookkeeping, transcribed: history = per-iteration loss vectors."""
    best_loss, best_iter, num_drop, lr, stopped = 1e20, 0, 0, base_lr, False
    out = []
    for it, losses in enumerate(history):
        if stopped:
            out.append(dict(best_loss=best_loss, best_iter=best_iter, num_loss_drop=num_drop, lr=lr, stopped=True))
            continue
        rebuild = False
        if it in drops:
            rebuild = True
        else:
            loss_sum = np.float32(0)
            for v in losses:
                loss_sum = np.float32(loss_sum + np.float32(v))
            did = False
            if loss_sum < best_loss:
                best_loss, best_iter = float(loss_sum), it
            elif it - best_iter >= delay:
                did = True
            if auto_stop:
                rebuild = did
        if rebuild:
            num_drop += 1
            if num_drop > max_drops:
                stopped = True
            else:
                best_iter, best_loss = it, 1e20
                lr = base_lr / 10 ** num_drop
        out.append(dict(best_loss=best_loss, best_iter=best_iter, num_loss_drop=num_drop, lr=lr, stopped=stopped))
    return out


def test_device_checkdrop_equals_the_reference_policy():
    from test_pipeline_gpu import build
    vq, clip, eng, prompts, z = build(cutn=8, seed=13)
    base_lr, delay, max_drops, drops = 0.08, 2, 2, [9]
    eng.set_schedule(base_lr, iter_drop_delay=delay, max_loss_drops=max_drops, auto_stop=True, drops=drops)
    zc = z.clone().cuda()
    history, records, zs = [], [], []
    for it in range(40):
        eng.iterate(zc, 123.0, it)            # the lr argument is ignored in managed mode
        eng.sync()
        rec = eng.poll_status()
        assert rec is not None and rec["iter"] == it
        history.append(rec["losses"].copy())
        records.append(rec)
        zs.append(zc.clone())
    want = host_policy(history, base_lr, delay, max_drops, True, drops)
    for it, (r, w) in enumerate(zip(records, want)):
        assert r["num_loss_drop"] == w["num_loss_drop"] and r["stopped"] == w["stopped"] and r["best_iter"] == w["best_iter"], (it, r, w)
        assert abs(r["best_loss"] - w["best_loss"]) <= 1e-6 * max(1.0, abs(w["best_loss"])) or (r["best_loss"] > 1e19 and w["best_loss"] > 1e19)
        assert abs(r["lr"] - w["lr"]) <= 1e-7
    n_stop = [i for i, w in enumerate(want) if w["stopped"]]
    assert want[-1]["num_loss_drop"] >= 2, "the scripted run must exercise at least two drops"
    assert n_stop, "the scripted run must reach the stop"
    first = n_stop[0]
    # the stopping iteration still took its step (opt.step() precedes the return False, pixray.py:1484-1506) ...
    assert not torch.equal(zs[first], zs[first - 1])
    # ... and everything enqueued afterwards leaves z untouched
    for k in range(first + 1, 40):
        assert torch.equal(zs[k], zs[first])
    # best_z is the latent BEFORE the step of the best iteration (get_z_copy at checkdrop time, pixray.py:1104)
    bi = records[first]["best_iter"]
    if records[first]["best_loss"] < 1e19 and bi > 0:
        best_z = eng.debug_read("best_z", z.shape)
        assert torch.equal(best_z, zs[bi - 1])


def test_managed_adam_equals_the_plain_step():
    """Same gradient, same state: the managed kernel's Adam + clip_z update equals pxr_step's (which the pipeline tests
    pin to torch.optim.Adam semantics), including the fresh optimiser after a scheduled drop."""
    from test_pipeline_gpu import build, random_transforms
    vq, clip, eng_a, prompts, z = build(cutn=8, seed=17)
    _, _, eng_b, _, _ = build(cutn=8, seed=17)
    T = random_transforms(8, 224, 3)
    eng_a.set_schedule(0.05, iter_drop_delay=12, max_loss_drops=1, auto_stop=False, drops=[2])
    za, zb = z.clone().cuda(), z.clone().cuda()
    lr = 0.05
    for it in range(5):
        p = dict(transforms=T, zoom_padding=it % 2, fill=0.4)
        eng_a.iterate(za, 0.0, it, params=p)
        # plain engine: same forward / backward, host-driven step with the lr the policy dictates
        eng_b.iterate(zb, lr, it, params=p)
        if it == 2:
            lr = 0.05 / 10
            eng_b.reset_optimizer()
        eng_a.sync()
        eng_b.sync()
        ga, gb = eng_a.debug_read("z_grad", z.shape), eng_b.debug_read("z_grad", z.shape)
        gmax = gb.abs().max().item()
        assert (ga - gb).abs().max().item() <= 5e-3 * gmax          # atomics in the cutout backward, re-quantised by the fp16 decoder backward
        solid = gb.abs() > 5e-2 * gmax                                # Adam's first step is sign-like: skip noise-level entries
        assert ((za - zb).abs() * solid).max().item() < 2e-4, it   # lr * (fp16-level gradient noise)
        zb.copy_(za)                                                  # keep both engines on the same trajectory


def test_batches_accumulate_the_gradient_of_fresh_passes():
    from test_pipeline_gpu import build
    vq, clip, eng, prompts, z = build(cutn=8, seed=19)
    it, lr = 3, 0.05
    grads = []
    for b in range(3):
        key = it if b == 0 else it + b * (1 << 20)
        eng.synth(z)
        eng.make_cutouts(None, use_engine_rng=True, it=key)
        eng.encode_image(0)
        eng.prompt_loss(0)
        grads.append(eng.backward().clone())
    assert (grads[0] - grads[1]).abs().max() > 1e-3 * grads[0].abs().max()      # the passes draw different augmentations
    want = grads[0] + grads[1] + grads[2]
    eng.set_batches(3)
    eng.reset_optimizer()
    zc = z.clone().cuda()
    losses = np.zeros(2, dtype=np.float32)
    eng.iterate(zc, lr, it, losses_out=losses)
    acc = eng.debug_read("z_grad_acc", z.shape)
    assert (acc - want).abs().max().item() <= 5e-3 * want.abs().max().item()
    # one Adam step on the accumulated gradient
    from oracle import ref_path as R
    zmin, zmax = R.vqgan_z_bounds(vq)
    z_next = torch.maximum(torch.minimum(R.AdamState(z).step(z, acc.cpu(), lr), zmax), zmin)
    solid = acc.cpu().abs() > 5e-2 * acc.abs().max().item()
    assert ((zc.cpu() - z_next).abs() * solid).max().item() < 2e-5
    # the reported losses are the first pass's (`if i == 0`, pixray.py:1466)
    eng.set_batches(1)
    eng.synth(z)
    eng.make_cutouts(None, use_engine_rng=True, it=it)
    eng.encode_image(0)
    first = eng.prompt_loss(0).cpu().numpy()
    assert np.abs(first - losses).max() < 1e-5


@pytest.mark.parametrize("device_checkdrop", [True, False])
def test_api_run_with_auto_stop_and_batches(device_checkdrop):
    """pixray.run() with the presets that need this control flow: auto_stop (drops used up -> the run ends early) and
    batches = 2; both the device-managed loop and the reference's host-side flow."""
    from pixray_b200 import api
    api.run("a cat", "vqgan", size=[128, 128], clip_models="ViT-B/16", iterations=60, num_cuts=8, batches=2, outdir="",
            vector_prompts="none", b200_allow_synthetic=True, seed="5", learning_rate_drops=[10], auto_stop=True,
            learning_rate=0.5, b200_device_checkdrop=device_checkdrop)
    st = api._state
    img = api.get_image()
    assert st.num_loss_drop >= 1 and (img is None or torch.isfinite(img).all())
    assert st.cur_iteration <= 60 and np.isfinite(st.losses).all()


def test_checkpoint_resume_follows_the_uninterrupted_run():
    """pxr_save_state / pxr_load_state (z, Adam m / v / step, drop bookkeeping): a session restored into a fresh engine and
    continued equals the uninterrupted one (engine-drawn augmentations are keyed by (seed, iteration)); the last bits of
    the gradients differ run to run (atomics), hence a small tolerance instead of bit equality."""
    from test_pipeline_gpu import build
    vq, clip, eng_a, prompts, z = build(cutn=8, seed=23)
    _, _, eng_b, _, _ = build(cutn=8, seed=23)
    for e in (eng_a, eng_b):
        e.set_schedule(0.05, iter_drop_delay=12, max_loss_drops=2, auto_stop=False, drops=[1])
    za = z.clone().cuda()
    for it in range(3):
        eng_a.iterate(za, 0.0, it)
    eng_a.sync()
    blob = eng_a.save_state()
    for it in range(3, 7):
        eng_a.iterate(za, 0.0, it)
    eng_a.sync()
    eng_b.load_state(blob)
    zb = eng_b.read_z()
    for it in range(3, 7):
        eng_b.iterate(zb, 0.0, it)
    eng_b.sync()
    ra, rb = eng_a.poll_status(), eng_b.poll_status()
    assert ra["num_loss_drop"] == rb["num_loss_drop"] == 1 and abs(ra["lr"] - rb["lr"]) < 1e-9
    moved = (za - z.cuda()).abs().max().item()
    diff = (za - zb).abs()
    # last-bit gradient noise can flip a VQ code or a noise-level Adam sign somewhere: the bulk must agree tightly
    assert moved > 0.02 and diff.median().item() < 1e-3 * moved and (diff > 0.1 * moved).float().mean().item() < 0.02
    with pytest.raises(E.EngineError):
        eng_b.load_state(b"\\0" * len(blob))
