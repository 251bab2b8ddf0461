"""CPU: the C-ABI library loads and exports every symbol the headers declare (no compute calls -- there is no GPU
here), the ctypes table matches the headers, and the host-side logic (schedules, Generator kwargs merging, error
behaviour, LIM coefficient table, architecture walk) agrees with the oracle / golden vectors."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def header_symbols():
    syms = {}
    for h in ("dlpm_b200.h", "dlpm_b200_unet.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b(int|int64_t|const char\*)\s+(dlpm_b200_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
            args = [a.strip() for a in m.group(3).split(",") if a.strip() and a.strip() != "void"]
            syms[m.group(2)] = len(args)
    return syms


def test_library_exports_every_declared_symbol():
    from dlpm_b200 import _lib, _unet_lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 37
    for new in ("dlpm_b200_graph_sample", "dlpm_b200_graph_sample_stats", "dlpm_b200_reverse_step_post", "dlpm_b200_dlim_step_post",
                "dlpm_b200_lim_step_post", "dlpm_b200_set_counter", "dlpm_b200_philox_rounds"):
        assert new in syms, new
    lib.dlpm_b200_philox_rounds.restype = ctypes.c_int
    assert lib.dlpm_b200_philox_rounds() in (7, 10)
    for name in syms:
        assert hasattr(lib, name), "missing export: " + name
    lib.dlpm_b200_abi_version.restype = ctypes.c_int
    assert lib.dlpm_b200_abi_version() == 1
    # the ctypes signature table covers the headers and agrees on arity
    table = dict(_lib.SIGNATURES)
    table.update(_unet_lib.UNET_SIGNATURES)
    for name, nargs in syms.items():
        if name in ("dlpm_b200_abi_version", "dlpm_b200_last_error", "dlpm_b200_unet_workspace_bytes", "dlpm_b200_philox_rounds"):
            continue
        assert name in table, "no ctypes signature for " + name
        assert len(table[name]) == nargs, (name, len(table[name]), nargs)


def test_no_cpu_fallback_and_loud_failures():
    import dlpm_b200
    from dlpm_b200 import _lib
    with pytest.raises(_lib.DlpmB200Error, match="CUDA devices only"):
        dlpm_b200.gen_skewed_levy(1.7, (4, 4), device="cpu")
    with pytest.raises(Exception, match="Wrong value of alpha"):
        dlpm_b200.gen_skewed_levy(0.0, (4, 4), device="cpu")
    with pytest.raises(_lib.DlpmB200Error):
        _lib.ptr(torch.zeros(3))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.DlpmB200Error):
            _lib.require_cuda("cuda")
    # the product never imports the oracle
    import subprocess
    import sys
    code = "import sys; import dlpm_b200, dlpm_b200.score_nets, dlpm_b200._unet_lib; assert not any(m.split('.')[0]=='oracle' for m in sys.modules)"
    assert subprocess.run([sys.executable, "-c", code], cwd=ROOT).returncode == 0
    for base, _, files in os.walk(os.path.join(ROOT, "dlpm_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(base, f)).read().replace("oracle/", ""), f


def test_host_schedule_bit_exact_vs_reference_golden():
    from dlpm_b200.methods.dlpm import DLPM
    g = load_golden("schedule")
    for key in g.files:
        a, T, kind = key.split("_")
        alpha, T = float(a[1:]), int(T[1:])
        d = DLPM(alpha, "cpu", T, scale="scale_exploding") if kind == "exploding" else DLPM(alpha, "cpu", T, time_spacing=kind)
        got = torch.stack([d.gammas, d.bargammas, d.sigmas, d.barsigmas]).numpy()
        assert np.array_equal(got, g[key], equal_nan=True), key
        assert np.array_equal(d.sched.numpy(), g[key].T, equal_nan=True)
    d = DLPM(1.7, "cpu", 50)
    d.rescale_diffusion(10)
    assert d.gammas.shape[0] == 10 and d.sched.shape == (10, 4)
    with pytest.raises(AssertionError):
        d.rescale_diffusion(10.0)


def test_generator_kwargs_semantics():
    from dlpm_b200 import Generator
    gen = Generator("skewed_levy", alpha=1.7, device="cpu", isotropic=True, clamp_a=None)
    gen.setParams(clamp_a=20)
    assert gen.kwargs["clamp_a"] == 20 and gen.kwargs["alpha"] == 1.7
    with pytest.raises(Exception, match="Given void parameters"):
        gen.setParams()
    with pytest.raises(Exception, match="Unknown distribution"):
        Generator("gmm_2")
    with pytest.raises(Exception):  # reaches the kernel wrapper, which refuses non-CUDA devices
        gen.generate(size=(4, 2))
    assert "clamp_a" in str(gen.getSignature())


def test_glp_constructor_contract():
    from dlpm_b200 import GenerativeLevyProcess, ModelMeanType
    glp = GenerativeLevyProcess(1.7, "cpu", 100, rescale_timesteps=True)
    assert glp.dlpm.gammas.shape == (100,) and glp.reverse_steps == 100 and not glp.LIM
    assert torch.equal(glp.get_timesteps(5), torch.arange(5, dtype=torch.float32))
    assert torch.allclose(glp._scale_timesteps(torch.tensor([50])), torch.tensor([0.5]))
    with pytest.raises(AssertionError):
        GenerativeLevyProcess(1.7, "cpu", 100, model_mean_type=ModelMeanType.START_X)
    with pytest.raises(AssertionError):
        GenerativeLevyProcess(1.7, "cpu", 100, LIM=True, rescale_timesteps=False)
    lim = GenerativeLevyProcess(1.7, "cpu", 100, LIM=True, rescale_timesteps=True)
    assert abs(lim.sde.T - 0.9946) < 1e-9
    with pytest.raises(AssertionError, match="time spacing"):
        glp.sample({"default": None}, [2, 1, 2], 100, time_spacing="quadratic")


def test_lim_table_matches_oracle():
    from dlpm_b200.methods.lim import VPSDE, lim_step_table
    from oracle import process
    for ode in (False, True):
        ts, coef = lim_step_table(VPSDE(1.7), 25, ode)
        sde = process.VPSDE(1.7)
        sc, a, cs, cn = process.lim_coefficients(sde, ts[:-1], ts[1:], ode)
        assert torch.equal(coef[:, 0], sc) and torch.equal(coef[:, 1], a) and torch.equal(coef[:, 2], cs)
        if not ode:
            assert torch.equal(coef[:, 3], cn)


def test_unet_program_matches_oracle_block_plan():
    from dlpm_b200.score_nets import OP_ATTN, OP_CONV, OP_CONV_IN, OP_GN, OP_SPLIT, UNetModel
    from oracle import nets
    for mc, attn in ((128, (16,)), (32, (2, 4))):
        m = UNetModel(3, mc, 3, 2, attn, channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
        prog = m.build_program(32, 32, fuse_gn=False, fuse_gne=False)
        inp, out = nets.unet_block_plan(mc, (1, 2, 2, 2), 2, attn)
        n_res = sum(l.count("res") for l in inp + out) + 2
        n_attn = sum(l.count("attn") for l in inp + out) + 1
        ops = [o[0] for o in prog["ops"]]
        assert ops.count(OP_ATTN) == n_attn
        assert ops.count(OP_GN) == 2 * n_res + n_attn + 1
        # GroupNorms attached to their producing convolutions (default): every GroupNorm is either an op or a fused target
        # (a concatenation's GroupNorm counts once but is carried by BOTH producers)
        m.fuse_groupnorm_max_pixels = 1024
        fused = m.build_program(32, 32, fuse_gn=True, fuse_gne=False)
        fops = [o[0] for o in fused["ops"]]
        dsts = {o[24 + 8 * k] for o in fused["ops"] if o[0] == OP_CONV for k in (0, 1) if o[24 + 8 * k] >= 0}
        assert fops.count(OP_GN) + len(dsts) == 2 * n_res + n_attn + 1
        if mc == 128:
            # benchmark net: only the GroupNorms fed by a folded upsample conv (3) or with 12-channel groups straddling the
            # 256 | 128 concatenation (1) stay separate launches
            assert fops.count(OP_GN) == 4, fops.count(OP_GN)
        else:
            assert len(dsts) == 0  # 32 / 64 channels: groups smaller than a channel quad
        assert [o for o in fops if o != OP_GN] == [o for o in ops if o != OP_GN]
        if mc == 128:  # restricted to the 8x8 and 4x4 maps (whole samples inside one tile)
            m.fuse_groupnorm_max_pixels = 64
            dflt = m.build_program(32, 32, fuse_gn=True, fuse_gne=False)
            d_dsts = {o[24 + 8 * k] for o in dflt["ops"] if o[0] == OP_CONV for k in (0, 1) if o[24 + 8 * k] >= 0}
            assert all(o[8] * o[9] // (o[13] * o[13]) <= 64 for o in dflt["ops"] if o[0] == OP_CONV and o[24] >= 0)
            assert len(d_dsts) == 24 and [o[0] for o in dflt["ops"]].count(OP_GN) + len(d_dsts) == 2 * n_res + n_attn + 1, len(d_dsts)
        if mc == 128:  # the default engine: GroupNorm in the epilogue (GNE) on the 16x16 maps (one target per 256-channel conv) and 4x4 maps
            gne = m.build_program(32, 32, fuse_gn=False)
            g_ops = [o for o in gne["ops"] if o[0] == OP_CONV and o[24] >= 0]
            hw = lambda o: o[8] * o[9] // (o[13] * o[13])
            assert all(hw(o) in (16, 64, 256) for o in g_ops) and all(o[24 + 8] < 0 for o in g_ops if hw(o) == 256)
            g_dsts = {o[24 + 8 * k] for o in g_ops for k in (0, 1) if o[24 + 8 * k] >= 0}
            assert [o[0] for o in gne["ops"]].count(OP_GN) + len(g_dsts) == 2 * n_res + n_attn + 1
            assert not any(o[0] == OP_GN and o[6] == 16 for o in gne["ops"])  # no GroupNorm launch left on the 4x4 maps
            assert sum(1 for o in gne["ops"] if o[0] == OP_GN and o[6] == 64) == 1  # 8x8: only the one fed by the folded upsample conv
            # conv1-type outputs (only reader = the fused GroupNorm) carry the "raw output unused" flag
            assert sum(1 for o in g_ops if hw(o) == 256) == 7 and sum(1 for o in g_ops if hw(o) == 256 and o[31] & 2) == 5
        n_up = sum(l.count("up") for l in inp + out)
        n_down = sum(l.count("down") for l in inp + out)
        # an upsample+conv = ONE parity-batched launch; the input conv = bf16 split + ONE tensor-core conv
        assert ops.count(OP_SPLIT) == 1 and ops.count(OP_CONV_IN) == 0
        assert ops.count(OP_CONV) == 2 * n_res + 2 * n_attn + n_down + n_up + 2
        assert prog["header"][7] == sum(2 * b.out_channels for b in m.modules() if hasattr(b, "emb_layers"))
        assert prog["wb"].dtype == torch.bfloat16 and prog["wb"].numel() % 64 == 0
    # state_dict key set equals the oracle's expectations (keys it reads exist)
    sd = m.state_dict()
    for k in ("time_embed.0.weight", "input_blocks.0.0.weight", "middle_block.1.qkv.weight", "out.2.bias",
              "input_blocks.3.0.op.weight", "output_blocks.2.1.conv.weight", "output_blocks.0.0.skip_connection.weight"):
        assert k in sd, k


def test_mlp_packing_layout():
    from dlpm_b200.score_nets import MLPModel
    p = {"data": {"nfeatures": 2}, "method": "dlpm", "dlpm": {"isotropic": True}, "device": "cpu",
         "model": dict(use_a_t=False, no_a=True, a_pos_emb=False, a_emb_size=32, time_emb_type="learnable", time_emb_size=32,
                       nblocks=4, nunits=64, skip_connection=True, group_norm=True, dropout_rate=0.0, learn_variance=False)}
    m = MLPModel(p)
    w = m.packed_weights()
    E, U, F, NB = 32, 64, 2, 5
    main0 = 3 * E + E * E + NB * (E * U + U)
    assert w.numel() == main0 + F * U + 3 * U + NB * (2 * U * U + 6 * U) + F * U + 4
    assert torch.equal(w[main0:main0 + F * U].reshape(F, U), m.linear_in.weight.t())
    with pytest.raises(NotImplementedError):
        bad = dict(p)
        bad["model"] = dict(p["model"], time_emb_type="sinusoidal")
        MLPModel(bad)


def test_unet_program_interpreted_full_width_fused_groupnorm():
    """Benchmark-width op list with the GroupNorms attached to their producing convolutions (fused targets, both halves of
    the skip concatenations) against the reference output."""
    from program_interpreter import interpret
    from dlpm_b200.init_utils import randomize_parameters_
    from dlpm_b200.score_nets import UNetModel
    g = load_golden("unet_cifar_full")
    m = UNetModel(3, 128, 3, 2, (16,), channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
    randomize_parameters_(m, 21)
    want = torch.from_numpy(g["y"])
    for fuse, max_px in ((True, 64), (True, 1024), (False, 64)):
        m.fuse_groupnorm_max_pixels = max_px
        y, _ = interpret(m.build_program(32, 32, fuse_gn=fuse), torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), 128)
        err = float((y - want).abs().max() / want.abs().max())
        assert err < 6e-3, (fuse, max_px, err)


def test_unet_program_interpreted_cifar10_yml_architecture():
    """cifar10.yml:51-59 (``attn_resolutions: [4, 8, 16]``: AttentionBlocks after every ResBlock of the 8x8 and 4x4 levels): the op
    list of every GroupNorm strategy against the reference output, batch-constant and per-sample t."""
    from program_interpreter import interpret
    from dlpm_b200.init_utils import parameter_checksum, randomize_parameters_
    from dlpm_b200.score_nets import OP_ATTN, UNetModel
    g = load_golden("unet_cifar10_attn")
    m = UNetModel(3, 128, 3, 2, (4, 8, 16), channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
    randomize_parameters_(m, 21)
    assert abs(parameter_checksum(m) - float(g["weight_checksum"])) < 1e-6 * abs(float(g["weight_checksum"]))
    assert sum(p.numel() for p in m.parameters()) == int(g["nparams"])
    x = torch.from_numpy(g["x"])
    for kw in (dict(), dict(fuse_gn=False, fuse_gne=False), dict(fuse_gn=True, fuse_gne=False)):
        prog = m.build_program(32, 32, **kw)
        assert [o[0] for o in prog["ops"]].count(OP_ATTN) == 11  # 2 + 2 (input blocks), 1 (middle), 3 + 3 (output blocks)
        for tk, yk in (("t", "y"), ("t2", "y2")):
            y, _ = interpret(prog, x, torch.from_numpy(g[tk]), 128)
            want = torch.from_numpy(g[yk])
            err = float((y - want).abs().max() / want.abs().max())
            assert err < 6e-3, (kw, tk, err)


@pytest.mark.parametrize("name", ["mnist", "cifar_half"])
def test_unet_program_interpreted_matches_reference_golden(name):
    """The op list + packed weights, executed by a plain-PyTorch interpreter (fp32 except bf16 conv weights),
    reproduce the reference UNet output: validates the architecture walk / packing the C++ engine consumes."""
    from program_interpreter import interpret
    from test_oracle_golden import UNET_CFGS
    from dlpm_b200.init_utils import randomize_parameters_
    from dlpm_b200.score_nets import UNetModel
    g = load_golden("unet_" + name)
    c = UNET_CFGS[name]
    m = UNetModel(in_channels=c["in_ch"], model_channels=c["model_channels"], out_channels=c["in_ch"],
                  num_res_blocks=c["num_res_blocks"], attention_resolutions=c["attention_resolutions"],
                  channel_mult=c["channel_mult"], num_heads=c["num_heads"], use_scale_shift_norm=True)
    randomize_parameters_(m, 21)
    prog = m.build_program(32, 32)
    x = torch.from_numpy(g["fwd/x"])
    for tk, yk in (("fwd/t", "fwd/y"), ("fwd2/t", "fwd2/y")):
        y, _ = interpret(prog, x, torch.from_numpy(g[tk]), c["model_channels"])
        want = torch.from_numpy(g[yk])
        err = float((y - want).abs().max() / want.abs().max())
        assert err < 6e-3, (name, tk, err)  # only the conv weights are bf16-rounded here
