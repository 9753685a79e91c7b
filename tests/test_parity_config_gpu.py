"""Parity at the BENCHMARKED configurations (BASELINE.json configs[1] and configs[2]): the CUDA train step against the
oracle at Stage-I batch 128 and Stage-II batch 64, in bf16 (tensor-core) and fp32 (validation) mode.

Checker: the oracle itself (oracle/stackgan_oracle.py, pinned to the unmodified reference by tests/golden) run on the
GPU in fp64 -- tests/gpu_oracle.py.  Every quantity the reference's step produces is compared: fake images, the four
critic scores, the gradient penalty, both losses, and -- per parameter tensor -- the gradients each optimizer sees at
its step (for Stage-II's generator that is the sum over the five critic backward passes plus lossG's,
stage_2_train_fn.py:131,154,163-168).  The critic is re-synchronised to the fp64 trajectory before each iteration
(see tests/test_stage1_gpu.py) so that iterations 2..5 compare kernels, not Adam's sign(g) amplification.

Three numbers per tensor go into the report (gpurun_out/parity_<stage>_<mode>_B<batch>.txt, copied to profiles/):
  rel_l2      ||got - ref|| / ||ref||
  yardstick   the same for the *reference precision*: the fp32 oracle (fp32 mode) or the engine's dataflow evaluated
              exactly with ideal bf16 storage rounding (bf16 mode, tests/emu_ops.py on the GPU) -- what no
              implementation storing bf16 activations can beat
  worst/bound max |got-ref| / (rtol |ref| + atol max|ref| + 3 yardstick_max)   with BASELINE.json's rtol / atol

  vs_yard     ||got - yardstick implementation|| / ||ref||

A gradient bound cannot go vacuous silently: where the yardstick is claimed small (Stage-I in both modes, Stage-II in
fp32) the test ASSERTS rel_l2(yardstick) < YARD_L2 for every gradient tensor before `3 x yardstick` counts as slack, and
independently requires rel_l2(got) <= L2_FACTOR x rel_l2(yardstick) + L2_FLOOR (both taken without the max(1, numel/1000)
largest-error elements -- mask flips, which the elementwise criterion bounds separately by OUTLIER_MAX).  Stage-II in bf16 is the one case where
the yardstick is NOT small (measured: bf16 rounding of the FORWARD activations alone -- backward tensors kept exact --
leaves the generator's gradients 0.3-0.9 away from fp64 in relative L2, the fp32 reference itself is 1e-2 away;
profiles/exp_bf16_sensitivity_r2.txt): there the report says so, fp32 mode at the same batch is the numerical gate, and
the bf16 run must (a) be no worse than the ideal-bf16 model tensor by tensor (factor 1.5) and (b) keep each optimizer's
whole gradient aligned with the oracle's (cosine) as well as the ideal model does -- neither of which zeros or a wrong
kernel can satisfy.
"""
import os

import pytest
import torch

import gpu_oracle as GO
from test_stage1_gpu import _modules as _modules1, _ref_table as _ref_table1, _run_teacher_forced as _run_tf1

pytestmark = pytest.mark.gpu

TOL = {"fp32": (1e-4, 1e-4), "bf16": (2e-2, 1e-3)}      # BASELINE.json: rtol / atol per mode
YARD_L2 = 0.2          # where the yardstick is claimed small, its relative L2 must stay below this for every gradient tensor
L2_FACTOR = {"fp32": 3.0, "bf16": 1.5}      # rel_l2(got) <= L2_FACTOR * rel_l2(yardstick) + L2_FLOOR
# fp32 floor: two or three mask flips in the 512-element BatchNorm gradients of the critic's last layer are 1.4e-3 of relative L2
# by themselves, and whether they land in the CUDA run or in the fp32 yardstick run changes from run to run (three runs of
# the same build, round 2: pass, 1.38e-3 on it0 down_sampler.6.1.bias against a yardstick of 3.6e-7, pass) -- 5e-4 was at that
# noise level.  Round 1's allowance was 1e-2.
L2_FLOOR = {"fp32": 2e-3, "bf16": 2e-2}
# one (Leaky)ReLU mask flip of a pre-activation within an ulp of zero moves a gradient element by ~1e-3 of the tensor's max in
# EITHER implementation, and which element flips is chance: the fp32 oracle's own deviation from fp64 shows it on some
# tensors (yard/max up to 2e-3 below) and not on others (2e-7).  Gradient rows get this much absolute slack in fp32 mode.
FLIP_ATOL = {"fp32": 1e-3, "bf16": 0.0}
ZERO_RMS = {"fp32": 1e-7, "bf16": 1e-4}
COS_SLACK, COS_MIN = 0.03, 0.3
# gradient tensors: the elementwise bound must hold for all but max(1, numel/1000) elements -- or, for weight tensors, for all
# but max(1, 2 %) of the output-channel rows -- (mask flips, see FLIP_ATOL and compare()), and no element may exceed it by more
# than OUTLIER_MAX
OUTLIER_MAX = 10.0
REPORT_ONLY = os.environ.get("SG_PARITY_REPORT_ONLY") == "1"


def _is_grad(k):
    return "/" in k or k == "dtem"


def _group(k):
    """gradient tensors that one optimizer sees together, e.g. 'it0 dD2' / 'G dG2' / 'dCA'."""
    return k.split("/")[0]


def compare(mode, want, got, yard, tag, yard_small=True):
    """Returns (report lines, failures).

    ``yard_small=True``: the yardstick is claimed to be << 1 and the test asserts it (then `3 x yardstick` slack in the
    elementwise bound means something).  ``False`` (Stage-II in bf16, where forward bf16 rounding alone makes generator
    gradients 0.3-0.9 off in relative L2 for ANY implementation -- profiles/exp_bf16_sensitivity_r2.txt): the per-tensor
    L2 criterion against the yardstick still applies, and each optimizer's whole gradient must point the same way as the
    oracle's at least as well as the ideal-bf16 model's does (cosine), which zeros or garbage cannot."""
    import math
    rt, at = TOL[mode]
    lines, fails, grels = [], [], []
    dots = {}
    hdr = (f"{'tensor':58s} {'rel_l2':>10s} {'yardstick':>10s} {'worst/bound':>11s} {'max|ref|':>10s} {'yard/max':>9s} "
           f"{'vs_yard':>10s}")
    lines.append(hdr)
    for k, r in want.items():
        if r is None or k not in got:
            continue
        g = got[k].detach().double().cpu().reshape(-1)
        r = r.detach().double().cpu().reshape(-1)
        y = yard[k].detach().double().cpu().reshape(-1) if k in yard and yard[k] is not None else None
        isg = _is_grad(k)
        if isg and r.norm().item() / max(r.numel(), 1) ** 0.5 < 1e-12:
            # exactly zero in exact arithmetic (the critic's compress.* gradients: the real / mismatched / fake text terms
            # cancel, stage_1_train_fn.py:140-144) -- nothing to be relative to; hold the result to an absolute rms
            rms = g.norm().item() / max(g.numel(), 1) ** 0.5
            lines.append(f"{k:58s} zero in exact arithmetic; rms(got) {rms:.3e}")
            if rms > ZERO_RMS[mode]:
                fails.append(f"[{tag}] {k}: reference is exactly zero, got rms {rms:.3e} > {ZERO_RMS[mode]}")
            continue
        scale = max(r.abs().max().item(), 1e-30)
        rn = max(r.norm().item(), 1e-30)
        rel = (g - r).norm().item() / rn
        yrel = (y - r).norm().item() / rn if y is not None else 0.0
        ymax = (y - r).abs().max().item() if y is not None else 0.0
        vs = (g - y).norm().item() / rn if y is not None else 0.0        # CUDA result against the yardstick implementation
        bound = rt * r.abs() + at * (scale if isg else 1.0) + 3.0 * ymax + (FLIP_ATOL[mode] * scale if isg else 0.0) + 1e-7
        ratio = (g - r).abs() / bound
        worst = ratio.max().item()
        if isg and worst > 1.0:
            allowed = max(1, r.numel() // 1000)
            kth = torch.topk(ratio, min(allowed + 1, ratio.numel())).values[-1].item()
            # ONE flipped mask at a layer's output moves a whole row of that layer's weight gradient (one output channel: all
            # its Ci*k*k elements, 2048 of the 524288 of a 128->256 k4 conv), so for weight tensors the flips are counted in
            # rows: at most max(1, 2 %) of the output channels may hold elements above the bound.  Anything broader than that
            # is caught by the L2 criterion, anything larger than OUTLIER_MAX x bound fails outright.
            shape = want[k].shape
            rows = shape[0] if len(shape) >= 2 else 0
            bad_rows = int((ratio.reshape(rows, -1).max(dim=1).values > 1.0).sum()) if rows else 0
            rows_ok = rows > 0 and bad_rows <= max(1, rows // 50)
            if (kth <= 1.0 or rows_ok) and worst <= OUTLIER_MAX:
                lines.append(f"{k:58s} (elementwise: {int((ratio > 1).sum())} of {r.numel()} elements above the bound"
                             + (f" in {bad_rows} of {rows} rows" if rows else "") + f", worst {worst:.2f})")
                worst_for_assert = min(kth, 1.0)
            else:
                worst_for_assert = worst
        else:
            worst_for_assert = worst
        lines.append(f"{k:58s} {rel:10.3e} {yrel:10.3e} {worst:11.3f} {scale:10.3e} {ymax / scale:9.2e} {vs:10.3e}")
        if isg:
            grels.append(rel)
            d = dots.setdefault(_group(k), [0.0] * 5)
            d[0] += (g * r).sum().item(); d[1] += (g * g).sum().item(); d[2] += (r * r).sum().item()
            if y is not None:
                d[3] += (y * r).sum().item(); d[4] += (y * y).sum().item()
            if yard_small and yrel > YARD_L2:
                fails.append(f"[{tag}] {k}: yardstick too coarse to judge with (rel_l2 {yrel:.2e} > {YARD_L2})")
            # the L2 criterion is taken over all but the `allowed` largest-error elements of BOTH sides (the same elements the
            # elementwise criterion sets aside as mask flips, bounded by OUTLIER_MAX there): one flipped element of a 512-element
            # BatchNorm gradient is 80 % of that tensor's squared error, and whether the flip lands in the CUDA run or in the
            # fp32 yardstick run is chance -- the untrimmed figure made this test pass or fail on the yardstick's own
            # run-to-run noise (3.49e-3 against 3 x 1.008e-3 + 5e-4 one run, 3 x 8.88e-4 + 5e-4 the next)
            allowed = max(1, r.numel() // 1000)
            def trimmed(e):
                if e.numel() <= allowed:
                    return 0.0
                e2 = e * e
                return math.sqrt(max(e2.sum().item() - torch.topk(e2, allowed).values.sum().item(), 0.0)) / rn
            rel_t, yrel_t = trimmed(g - r), (trimmed(y - r) if y is not None else 0.0)
            if rel_t > L2_FACTOR[mode] * yrel_t + L2_FLOOR[mode]:
                fails.append(f"[{tag}] {k}: rel_l2 {rel_t:.3e} (all elements: {rel:.3e}) > {L2_FACTOR[mode]} x yardstick "
                             f"{yrel_t:.3e} + {L2_FLOOR[mode]}  [both without their {allowed} largest-error element(s)]")
        if worst_for_assert > 1.0:
            fails.append(f"[{tag}] {k}: max err / bound = {worst:.3f} (rel_l2 {rel:.3e}, yardstick {yrel:.3e})")
    for grp, (gr, gg, rr, yr, yy) in dots.items():
        cg = gr / max(math.sqrt(gg * rr), 1e-300)
        cy = yr / max(math.sqrt(yy * rr), 1e-300) if yy > 0 else 1.0
        lines.append(f"# cosine with the oracle, all tensors of '{grp}': got {cg:.4f}   yardstick {cy:.4f}")
        if cg < cy - COS_SLACK or cg < COS_MIN:
            fails.append(f"[{tag}] '{grp}': cosine {cg:.4f} against the oracle (yardstick {cy:.4f}, floor {COS_MIN})")
    if grels:
        gm = math.exp(sum(math.log(max(v, 1e-30)) for v in grels) / len(grels))
        lines.append(f"# {tag}: {len(grels)} gradient tensors, geometric-mean rel_l2 {gm:.3e}, max {max(grels):.3e}; "
                     f"{len(fails)} failures")
    return lines, fails


def _dump(name, lines):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", name), "w") as f:
        f.write("\n".join(lines) + "\n")


# ------------------------------------------------------------------------------------------------ Stage-I, B=128
@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_stage1_batch128(mode):
    from imagegenerator_b200.ops import CudaOps
    from emu_ops import EmuOps
    B = 128
    b, ref = GO.stage1(B, torch.float64)
    want = _ref_table1(ref)
    if mode == "fp32":
        _, r32 = GO.stage1(B, torch.float32, force=ref["critic_before"])
        yard = _ref_table1(r32)
    else:
        yard = _run_tf1(EmuOps(torch.bfloat16, device="cuda"), b, ref)
    got = _run_tf1(CudaOps(mode), b, ref)
    torch.cuda.synchronize()
    lines, fails = compare(mode, want, got, yard, f"stage1 {mode} B{B}")
    _dump(f"parity_stage1_{mode}_B{B}.txt", lines)
    assert REPORT_ONLY or not fails, "\n".join(fails[:12])


# ------------------------------------------------------------------------------------------------ Stage-II
def _modules2():
    from imagegenerator_b200.con_augment import ConditioningAugmentation
    from imagegenerator_b200.discrminator_1 import StageIDiscriminator
    from imagegenerator_b200.discriminator_2 import StageIIDiscriminator
    from imagegenerator_b200.generator_1 import StageIGenerator
    from imagegenerator_b200.generator_2 import StageIIGenerator
    torch.manual_seed(42)
    return dict(ca1=ConditioningAugmentation(512, 256, 128), d1=StageIDiscriminator(512, 128), g1=StageIGenerator(128, 100),
                ca2=ConditioningAugmentation(512, 256, 128), d2=StageIIDiscriminator(512, 128), g2=StageIIGenerator())


def _run_tf2(ops, b, ref, g1_state=None):
    """One Stage-II outer step with the critic re-synchronised to ``ref``'s trajectory before every iteration."""
    from imagegenerator_b200.engine2 import Stage2Engine
    ms = _modules2()
    if g1_state is not None:
        ms["g1"].load_state_dict(g1_state)
    eng = Stage2Engine(ms["ca1"], ms["g1"], ms["ca2"], ms["d2"], ms["g2"], b["real"].shape[0], ops=ops)
    dv = lambda t: t.to(ops.device).to(ops.f32).contiguous()
    eng.load_batch(dv(b["real"]), dv(b["tem"]), dv(b["tem"][b["perm"]]))
    z, e1, e2, eg = dv(b["z"]), dv(b["eps_ca"]), dv(b["eps_ca2"]), dv(b["eps_gp"])
    c = lambda t: t.detach().double().cpu().clone()
    load = lambda m, sd: m.load_state_dict({k: v.float() for k, v in sd.items()})
    out = {}
    for it in range(5):
        if it > 0:
            load(ms["d2"], ref["critic_before"][it])
            eng.d.refresh_weights()
        eng.critic_iteration(z[it], e1[it], e2[it], eg[it])
        out[f"it{it} s_real"], out[f"it{it} s_mis"], out[f"it{it} s_fake"] = c(eng.d.score[0]), c(eng.d.score[1]), c(eng.d.score[2])
        out[f"it{it} gp"], out[f"it{it} loss_critic"] = c(eng.losses[1]), c(eng.losses[0])
        for k, v in ms["d2"].named_parameters():
            out[f"it{it} dD2/{k}"] = c(v.grad)
        if it == 0:
            out["it0 fake_64"] = c(eng.g1.out.permute(0, 3, 1, 2))
            out["it0 fake_256"] = c(eng.g2.out.permute(0, 3, 1, 2))
            eng.sync_grads()
            for k, v in ms["g2"].named_parameters():
                out[f"it0 dG2/{k}"] = c(v.grad)
            for k, v in ms["ca2"].named_parameters():
                out[f"it0 dCA2/{k}"] = c(v.grad)
    load(ms["d2"], ref["critic_before"][5])
    eng.d.refresh_weights()
    # the generator step with its optimizer steps held back, so that the gradients G2 / CA2 are stepped with can be read
    ops_step = eng.optimizer_step
    grads = {}

    def capture(fp):
        eng.side.join()
        if fp is eng.g2.fp:
            eng.g2.fold_grads()
            for k, v in ms["g2"].named_parameters():
                grads[f"G dG2/{k}"] = c(v.grad)
        else:
            for k, v in ms["ca2"].named_parameters():
                grads[f"G dCA2/{k}"] = c(v.grad)
        ops_step(fp)
    eng.optimizer_step = capture
    eng.generator_step()
    out.update(grads)
    out["G s_fake"], out["lossG"] = c(eng.d.score[2]), c(eng.losses[2])
    return out


def _ref_table2(ref):
    t = {}
    f = ref["first"]
    t["it0 fake_64"], t["it0 fake_256"] = f["fake_64"], f["fake"]
    for it in range(5):
        sc = ref["scores"][it]
        for k in ("s_real", "s_mis", "s_fake", "gp"):
            t[f"it{it} {k}"] = sc[k]
        t[f"it{it} loss_critic"] = ref["loss_critic"][it]
        for k, v in ref["critic_grads"][it].items():
            t[f"it{it} dD2/{k}"] = v
    for k, v in ref["g2_grads_it0"].items():
        t[f"it0 dG2/{k}"] = v
    for k, v in ref["ca2_grads_it0"].items():
        t[f"it0 dCA2/{k}"] = v
    t["G s_fake"], t["lossG"] = ref["s_gen"], ref["lossG"]
    for k, v in ref["g2_grads"].items():
        t[f"G dG2/{k}"] = v
    for k, v in ref["ca2_grads"].items():
        t[f"G dCA2/{k}"] = v
    return t


@pytest.mark.parametrize("mode,B,state", [("bf16", 64, "fresh"), ("fp32", 64, "fresh"), ("bf16", 64, "handoff")])
def test_stage2_config_batch(mode, B, state):
    """``state``: 'handoff' = gen_1 as a Stage-I run leaves it (BatchNorm running statistics converged, the situation
    stage_2_train_fn.py:65-72 sets up); 'fresh' = SURVEY.md section 8d's default, running statistics still (0, 1)."""
    from imagegenerator_b200.ops import CudaOps
    from emu_ops import EmuOps
    from oracle import stackgan_oracle as O
    g1_state = GO.trained_stats_g1(O.init_all(42)) if state == "handoff" else None
    b, ref = GO.stage2(B, torch.float64, g1_state=g1_state)
    want = {k: (v.detach().double().cpu() if torch.is_tensor(v) else v) for k, v in _ref_table2(ref).items()}
    force = ref["critic_before"]
    del ref
    torch.cuda.empty_cache()
    if mode == "fp32":
        _, r32 = GO.stage2(B, torch.float32, force=force, g1_state=g1_state)
        yard = {k: (v.detach().double().cpu() if torch.is_tensor(v) else v) for k, v in _ref_table2(r32).items()}
        del r32
    else:
        yard = _run_tf2(EmuOps(torch.bfloat16, device="cuda"), b, {"critic_before": force}, g1_state)
    torch.cuda.empty_cache()
    got = _run_tf2(CudaOps(mode), b, {"critic_before": force}, g1_state)
    torch.cuda.synchronize()
    lines, fails = compare(mode, want, got, yard, f"stage2 {mode} B{B} {state}", yard_small=(mode == "fp32"))
    _dump(f"parity_stage2_{mode}_B{B}_{state}.txt", lines)
    assert REPORT_ONLY or not fails, "\n".join(fails[:12])
