"""CPU: host-side logic of the product (no kernel launches): the C-ABI library loads and
exports every symbol the header declares, schedules / tables, error behaviour without a
GPU, sharding (incl. a world_size-2 gloo run), the UNet restatement."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, assert_golden

HEADER = os.path.join(ROOT, "include", "bndm_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bndm_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    from bndm_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", ROOT, "-j8"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bndm_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.bndm_version() == 2
    assert lib.bndm_last_error() is not None


def test_product_refuses_cpu_tensors():
    import bndm_b200 as bb
    x = torch.randn(1, 3, 64, 64)
    L = torch.eye(4096)
    with pytest.raises(bb.BndmError):
        bb.get_noise_v2(torch.device("cpu"), x, L, torch.ones(1), None, "gaussianBN", "train", True)
    with pytest.raises(bb.BndmError):
        bb.iadb_step(x, x, torch.ones(1))
    with pytest.raises(NotImplementedError):          # get_noise_recent.py:58-59
        bb.get_noise_v2(torch.device("cpu"), torch.randn(1, 3, 16, 16), L, None, None, "gaussian")
    with pytest.raises(NotImplementedError):          # 'uniform' never returns in the reference
        bb.get_noise_v2(torch.device("cpu"), x, L, None, None, "uniform")
    with pytest.raises(NotImplementedError):
        bb.get_noise_v2(torch.device("cpu"), x, L, None, None, "nope")


def test_fused_unet_gemm_paths_are_gated():
    """K9 / K10 / K11 are CUDA-only fast paths: CPU tensors and unsupported shapes never reach the library."""
    import torch
    from bndm_b200 import fused_unet as fu
    x = torch.randn(2, 128, 64, 64)
    assert not fu._use_shortcut_tc(x, None, torch.randn(128, 128, 1, 1))
    assert not fu._use_linear_tc(torch.randn(1024, 512), torch.randn(1536, 512))
    assert fu.conv_in3x3_nhwc(torch.randn(1, 8, 4, 4), torch.randn(128, 8, 3, 3)) is None          # not a conv_in shape
    assert fu.conv_in3x3_nhwc(torch.randn(1, 3, 4, 4), torch.randn(224, 3, 3, 3)) is None          # 256 % (224 / 4) != 0


def test_product_does_not_import_oracle():
    src_dir = os.path.join(ROOT, "bndm_b200")
    for fn in os.listdir(src_dir):
        if fn.endswith(".py"):
            assert "oracle" not in re.sub(r'""".*?"""', "", open(os.path.join(src_dir, fn)).read(), flags=re.S), fn


def test_schedule_functions_match_golden():
    from bndm_b200 import get_scheduler, get_scheduler_gamma
    g = load_golden("schedules")
    for key in g.files:
        parts = key.split("_")
        T = int(parts[-1][1:])
        x = torch.arange(0, T + 1).float()
        if parts[0] == "alpha":
            got = get_scheduler(x, "linear", T)
        else:
            p = (float(parts[2][3:]), float(parts[3][1:]), float(parts[4][1:]))
            got = get_scheduler_gamma(x, parts[1], p, T)
        assert_golden(got.numpy(), g[key], key)


@pytest.mark.parametrize("B,params", [(1, (1000.0, 0.0, 3.0)), (3, (0.2, 0.0, 3.0)), (37, (0.2, 0.0, 3.0))])
def test_iadb_table_rows_are_the_reference_differences(B, params):
    """Per step AND per sample: the reference evaluates the schedule on (B,)-shaped tensors, and
    torch's CPU sigmoid is not bit-identical between its vector body and its scalar tail."""
    from bndm_b200.schedules import iadb_table, latent_table
    from oracle.sampler import _coefficients
    T = 250
    table, first = iadb_table(T, "linear", "sigmoid", params, batch=B)
    assert table.shape == (T, B, 4) and first.shape == (B,) and bool((first == 1.0).all())
    for row, t in enumerate(reversed(range(T))):
        a_s, a_e, g_s, g_e = _coefficients(t, B, "cpu", T, "linear", "sigmoid", params)
        assert torch.equal(table[row, :, 0], a_s - a_e) and torch.equal(table[row, :, 1], g_s - g_e)
        assert torch.equal(table[row, :, 2], a_e)
    if params[0] == 1000.0:
        # tau=1000 makes d_gamma the cancellation-dominated constant 0.0039736032
        assert abs(float(table[0, 0, 1]) - 0.0039736032) < 1e-9
    assert abs(float(table[:, 0, 0].double().sum()) - 1.0) < 1e-6       # telescoping: sum d_alpha = 1
    lt, lfirst = latent_table(250, batch=B)
    assert lt.shape == (250, B, 4)
    assert bool((lfirst == 1.0).all()) and float(lt[0, 0, 0]) == np.float32(250 / 250 - 249 / 250)


def test_ddim_scheduler_tables_match_oracle():
    from bndm_b200 import DDIMScheduler
    from oracle.sampler import DDIMTables
    for n, eta in ((100, 0.0), (50, 1.0), (1000, 0.5)):
        s = DDIMScheduler()
        s.set_timesteps(n)
        o = DDIMTables()
        o.set_timesteps(n)
        assert list(map(int, s.timesteps)) == list(map(int, o.timesteps))
        tb = s.coefficient_table(eta)
        for i, t in enumerate(o.timesteps):
            want = [float(v) for v in o.coefficients(int(t), eta)]
            assert [float(v) for v in tb[i, :5]] == want
        assert float(tb[-1, 5]) == 0.0 and float(tb[0, 5]) == float(o.timesteps[1])
    with pytest.raises(ValueError):
        DDIMScheduler().step(torch.zeros(1), 0, torch.zeros(1))


def test_shard_bounds_cover_batch():
    from bndm_b200.dist import global_white_draw, shard_bounds
    for B, W in ((256, 8), (128, 8), (64, 1), (10, 4), (3, 8)):
        spans = [shard_bounds(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
    full = global_white_draw((8, 3, 4, 4), 0, 0, 1)
    parts = torch.cat([global_white_draw((8, 3, 4, 4), 0, r, 4) for r in range(4)])
    assert torch.equal(full, parts)
    np.random.seed(0)
    assert np.array_equal(full.numpy(), np.random.randn(8, 3, 4, 4).astype(np.float32))   # iadb_bn.py:75,761


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["BNDM_ROOT"])
from bndm_b200.dist import broadcast_L, gather_images, global_white_draw, shard_bounds, broadcast_module
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
r = dist.get_rank()
L = torch.arange(16.).reshape(4, 4) if r == 0 else torch.zeros(4, 4)
broadcast_L(L)
assert torch.equal(L, torch.arange(16.).reshape(4, 4))
lin = torch.nn.Linear(3, 3); broadcast_module(lin)
w = [torch.empty_like(lin.weight) for _ in range(2)]; dist.all_gather(w, lin.weight.data); assert torch.equal(w[0], w[1])
B = 5
mine = global_white_draw((B, 3, 2, 2), 7, r, 2)
lo, hi = shard_bounds(B, r, 2)
assert mine.shape[0] == hi - lo
out = gather_images(mine * 2, B)                      # "sampling" = independent per-sample work
full = global_white_draw((B, 3, 2, 2), 7, 0, 1) * 2
assert torch.equal(out, full), (out, full)
dist.destroy_process_group()
print("ok", r)
"""


def test_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port), BNDM_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()


def test_unet_restatement_shapes_and_keys():
    from bndm_b200.unet import count_forward_flops, get_latent_model, get_model
    m = get_model(3, 6, 64)
    assert abs(sum(p.numel() for p in m.parameters()) / 1e6 - 113.7) < 0.1
    assert abs(count_forward_flops(m, 64, 64) / 1e9 - 31.06) < 0.05
    keys = m.state_dict().keys()
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.0.resnets.0.norm1.weight",
              "down_blocks.0.downsamplers.0.conv.weight", "down_blocks.4.attentions.1.to_q.weight",
              "down_blocks.4.attentions.0.to_out.0.bias", "mid_block.attentions.0.group_norm.weight",
              "mid_block.resnets.1.time_emb_proj.weight", "up_blocks.1.attentions.2.to_v.weight",
              "up_blocks.0.upsamplers.0.conv.weight", "up_blocks.5.resnets.2.conv_shortcut.weight",
              "conv_norm_out.weight", "conv_out.bias"):
        assert k in keys, k
    assert not any(k.startswith("down_blocks.5.downsamplers") or k.startswith("up_blocks.5.upsamplers") for k in keys)
    small = get_latent_model(256, 8).eval()
    x = torch.randn(2, 4, 16, 16)
    with torch.no_grad():
        a = small(x, torch.tensor(0.5), return_dict=False)[0]
        b = small(x, torch.tensor([0.5, 0.5])).sample
        c = small(x, 0.5).sample
    assert a.shape == (2, 8, 16, 16) and torch.equal(a, b) and torch.equal(a, c)


def test_streamk_schedule_is_consistent():
    """The tcgen05 contraction's static work split (csrc/common.cuh StreamK): host-side check that
    every k-stage is covered once and the combine kernel reads exactly the partials the GEMM writes."""
    from bndm_b200 import _lib
    lib = _lib.load()
    for n_tiles in (32, 16):
        for dense in (0, 1):
            for n_colblk in (1, 2, 3, 7):
                for sms in (148, 132, 1, 2, 7, 33, 4096):
                    rc = lib.bndm_debug_streamk_check(n_tiles, dense, n_colblk, sms)
                    assert rc == 0, (n_tiles, dense, n_colblk, sms, lib.bndm_last_error())
                    rc = lib.bndm_debug_streamk_check_sub(n_tiles, dense, n_colblk, sms, 2)    # 64-k pipeline stages
                    assert rc == 0, (n_tiles, dense, n_colblk, sms, "sub=2", lib.bndm_last_error())


def test_gemv_row_schedule_covers_every_quad_once_and_is_balanced():
    """K1g's row ownership (csrc/noise_gemv.cu): every needed quad of 4 rows belongs to exactly one CTA slot, slots are
    sorted longest first, and no CTA streams more than 2.5 % above the mean at the full triangular size."""
    import ctypes as C
    from bndm_b200 import _lib
    lib = _lib.load()
    for variant in (0,):
        for res32 in (0, 1):
            for dense in (0, 1):
                for sms in (148, 132, 160):
                    worst, total = C.c_int(), C.c_int()
                    rc = lib.bndm_debug_gemv_schedule_check(res32, dense, sms, variant, C.byref(worst), C.byref(total))
                    assert rc == 0, (variant, res32, dense, sms, lib.bndm_last_error())
                    # longest-first dealing: never more than one longest quad above the mean ...
                    assert worst.value <= total.value / sms + 32, (variant, dense, sms, worst.value)
                    if not res32 and not dense and sms == 148:       # ... and within 2.5 % at the production shape
                        assert worst.value <= 1.025 * total.value / sms, (variant, worst.value, total.value)
    # 1024 quads do not fit 100 CTAs x 8 slots: reported, not silently truncated
    assert lib.bndm_debug_gemv_schedule_check(0, 0, 100, 0, None, None) == _lib.ERR_UNSUPPORTED


def test_fused_unet_and_training_front_end_refuse_cpu():
    """No CPU fallback anywhere in the product: the fused evaluator and the training-side entry raise."""
    import bndm_b200 as bb
    from bndm_b200.fused_unet import fuse_unet, groupnorm_silu_nhwc
    from bndm_b200.unet import get_latent_model
    model = get_latent_model(256, 8)
    with pytest.raises(bb.BndmError):
        fuse_unet(model)                                   # float32 but on the CPU
    with pytest.raises(TypeError):
        fuse_unet(torch.nn.Linear(2, 2))
    with pytest.raises(bb.BndmError):
        groupnorm_silu_nhwc(torch.randn(1, 128, 4, 4), torch.nn.GroupNorm(32, 128))
    with pytest.raises(bb.BndmError):
        bb.get_noise_train(torch.device("cpu"), torch.randn(2, 3, 64, 64), torch.eye(4096), torch.ones(2), torch.ones(2))


def test_iadb_snapshot_uint8_follows_the_reference_driver():
    """iadb_bn.py:796-802: clamp for the final image, min-max for snapshots, truncating uint8 cast."""
    from bndm_b200.io import iadb_snapshot_uint8
    torch.manual_seed(0)
    x = torch.randn(3, 8, 8) * 0.7
    fin = iadb_snapshot_uint8(x, final=True)
    want = (torch.clamp((x + 1) / 2.0, 0.0, 1.0).permute(1, 2, 0).numpy() * 255).astype(np.uint8)
    assert fin.dtype == np.uint8 and fin.shape == (8, 8, 3) and np.array_equal(fin, want)
    snap = iadb_snapshot_uint8(x, final=False)
    assert snap.min() == 0 and snap.max() == 255
    # truncation, not rounding: 0.999 * 255 = 254.7 -> 254
    assert iadb_snapshot_uint8(torch.full((1, 1, 1), 0.998), final=True)[0, 0, 0] == int((0.998 + 1) / 2 * 255)


def test_contraction_kernel_is_tcgen05_and_tma_in_sass():
    """Static check on the shipped library: every gemm_tc_kernel instance issues tcgen05.mma (UTCHMMA), reads
    TMEM (LDTM) and stages operands with bulk/tensor TMA (UBLKCP / UTMALDG); no legacy HMMA path anywhere."""
    import shutil
    import subprocess
    import sys
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_evidence.py")],
                         capture_output=True, text=True, check=True).stdout
    blocks = [b for b in out.split("\n\n") if b.startswith("bndm::gemm_tc_kernel")]
    assert len(blocks) >= 8
    for b in blocks:
        assert "UTCHMMA=" in b and "LDTM=" in b and ("UBLKCP=" in b or "UTMALDG=" in b), b
        assert "STACK:0" in b, "register spill in the contraction kernel:\n" + b
    assert "HMMA=" not in out.replace("UTCHMMA=", "")


def test_streaming_contraction_is_tma_fed_packed_fma_in_sass():
    """K1g (gemv_kernel): stage loads are one bulk copy (UBLKCP) + one tensor-map load (UTMALDG), the arithmetic is packed
    fp32 FMA (FFMA2: 8 per column and quad, in a few unrolled copies), no tensor-core instruction of any kind, and the
    cfg-1 instance spills at most a few registers."""
    import re
    import shutil
    import subprocess
    import sys
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_evidence.py")],
                         capture_output=True, text=True, check=True).stdout
    blocks = {b.split("\n")[0]: b for b in out.split("\n\n") if b.startswith("bndm::gemv_kernel")}
    assert len(blocks) == 4
    for nc in (4, 8, 12, 16):
        b = blocks[f"bndm::gemv_kernel<(int){nc}>"]
        n_ffma2 = int(re.search(r"FFMA2=(\d+)", b).group(1))
        assert "UBLKCP=" in b and "UTMALDG=" in b and n_ffma2 >= 8 * nc and n_ffma2 % (8 * nc) == 0, b
        assert "UTCHMMA" not in b and "HMMA" not in b, b
    assert int(re.search(r"STACK:(\d+)", blocks["bndm::gemv_kernel<(int)12>"]).group(1)) <= 32


def test_c_abi_rejects_bad_arguments_before_touching_the_device():
    """Argument validation of the C ABI (include/bndm_b200.h "Errors"): bad calls return a negative code and a
    message through bndm_last_error() without launching anything, so this runs without a GPU.  The UNSUPPORTED
    code is what the Python mirror turns into the reference's NotImplementedError (iadb_bn.py:331)."""
    import ctypes as C
    from bndm_b200 import _lib
    lib = _lib.load()
    fake = C.c_void_p(4096)                       # aligned, never dereferenced: every call below fails validation first

    def err():
        return lib.bndm_last_error().decode()

    h = C.c_void_p()
    assert lib.bndm_prepare_L(None, 4096, 12, None, C.byref(h)) == _lib.ERR_ARG and "null" in err()
    assert lib.bndm_prepare_L(fake, 1024, 12, None, C.byref(h)) == _lib.ERR_UNSUPPORTED and "4096" in err()
    assert lib.bndm_get_noise_f32(None, fake, None, fake, None, None, 4, 3, 64, 1, None) == _lib.ERR_ARG
    assert lib.bndm_reserve_columns(None, 12, None) == _lib.ERR_ARG
    assert lib.bndm_free_L(None) in (_lib.OK, _lib.ERR_ARG)
    # IADB update: null pointers, bad shapes, channel count that the reference rejects, missing coefficient vectors
    assert lib.bndm_iadb_step_f32(None, fake, fake, fake, None, 1, 3, 4096, 3, None) == _lib.ERR_ARG
    assert lib.bndm_iadb_step_f32(fake, fake, fake, fake, None, 0, 3, 4096, 3, None) == _lib.ERR_ARG
    assert lib.bndm_iadb_step_f32(fake, fake, fake, fake, fake, 1, 3, 4096, 5, None) == _lib.ERR_UNSUPPORTED
    assert "5 channels" in err()
    assert lib.bndm_iadb_step_f32(fake, fake, fake, fake, None, 1, 3, 4096, 6, None) == _lib.ERR_ARG
    assert "coefficient" in err()
    assert lib.bndm_iadb_step_sched_f32(fake, fake, fake, None, None, None, 1, 3, 4096, 6, None) == _lib.ERR_ARG
    assert lib.bndm_iadb_step_sched_dnhwc_f32(fake, fake, fake, fake, None, None, 1, 3, 4096, 6, None) == _lib.ERR_ARG
    assert lib.bndm_ddim_step_f32(fake, fake, None, None, None, None, None, 0, 0, 0, None) == _lib.ERR_ARG
    # UNet glue kernels
    assert lib.bndm_groupnorm_nhwc_f32(fake, None, 0, None, None, 0, fake, fake, None, fake, 1, 128, 16, 24,
                                       1e-5, 1, None) == _lib.ERR_UNSUPPORTED      # 128 % 24 != 0
    assert lib.bndm_groupnorm_nhwc_f32(fake, fake, 130, None, None, 0, fake, fake, None, fake, 1, 128, 16, 32,
                                       1e-5, 1, None) == _lib.ERR_ARG              # second source wider than C
    assert lib.bndm_groupnorm_nhwc_f32(C.c_void_p(4100), None, 0, None, None, 0, fake, fake, None, fake, 1, 128, 16, 32,
                                       1e-5, 1, None) == _lib.ERR_ARG and "aligned" in err()
    assert lib.bndm_add_bias_nhwc_f32(fake, None, None, fake, fake, fake, 130, 128, None) == _lib.ERR_ARG   # n % C
    assert lib.bndm_upsample2x_nhwc_f32(fake, fake, 1, 4, 4, 6, None) == _lib.ERR_ARG                       # C % 4
    assert lib.bndm_attention_small_f32(None, fake, 1, 16, 512, 8, None) == _lib.ERR_ARG
    # the tensor-core GEMMs of the fused UNet (K9, K10) and conv_in (K11): shapes and alignment are checked before any device work
    assert lib.bndm_linear_tc_f32(fake, fake, None, fake, 256, 512, 100, None) == _lib.ERR_UNSUPPORTED and "multiple of 32" in err()
    assert lib.bndm_linear_tc_f32(fake, None, None, fake, 256, 512, 128, None) == _lib.ERR_ARG
    assert lib.bndm_linear_tc_f32(C.c_void_p(4100), fake, None, fake, 256, 512, 128, None) == _lib.ERR_ARG and "aligned" in err()
    assert lib.bndm_shortcut_residual_tf32(fake, None, 128, 64, fake, fake, None, fake, 4096, 128, None) == _lib.ERR_ARG     # C2 > 0 without x2
    assert lib.bndm_shortcut_residual_tf32(fake, fake, 120, 64, fake, fake, None, fake, 4096, 128, None) == _lib.ERR_UNSUPPORTED
    assert lib.bndm_shortcut_residual_tf32(fake, fake, 128, 64, fake, None, None, fake, 4096, 128, None) == _lib.ERR_ARG
    assert lib.bndm_conv_in3x3_nhwc_f32(fake, fake, None, 1, 3, 8, 8, 128, None) == _lib.ERR_ARG
    assert lib.bndm_to_uint8_nhwc(fake, None, 1, 3, 64, 64, None) == _lib.ERR_ARG
    assert lib.bndm_white128_reinterpret_f32(fake, fake, 1, 3, None) == _lib.ERR_ARG and "in-place" in err()
    # the mapping the Python mirror applies
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_UNSUPPORTED, "x")
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_ARG, "x")
    with pytest.raises(_lib.BndmError):
        _lib.check(_lib.ERR_CUDA, "x")


def test_L_construction_from_a_covariance():
    """covariance -> Cholesky factor.  cholesky_L reproduces blue_noise_L from the same covariance, and
    the factor of an empirical covariance regenerates that covariance (small tile so it runs in a second)."""
    from bndm_b200 import synth
    sigma = synth.blue_noise_sigma()
    assert sigma.shape == (4096, 4096) and np.allclose(np.diag(sigma), 1.0)
    sub = torch.from_numpy(sigma[:512, :512].copy())              # leading principal block: still SPD
    L = synth.cholesky_L(sub)
    assert L.dtype == torch.float32 and torch.equal(L, torch.tril(L))
    want = np.linalg.cholesky(sigma[:512, :512]).astype(np.float32)
    np.testing.assert_allclose(L.numpy(), want, rtol=0, atol=2e-6)
    g = torch.Generator().manual_seed(0)
    fields = (torch.randn(4000, 64, generator=g, dtype=torch.float64) @ L[:64, :64].double().T).reshape(4000, 8, 8)
    cov = synth.empirical_covariance(fields)
    assert cov.shape == (64, 64) and cov.dtype == torch.float64
    np.testing.assert_allclose(cov.numpy(), sigma[:64, :64], atol=0.08)
    L2 = synth.cholesky_L(cov, jitter=1e-9)
    np.testing.assert_allclose((L2.double() @ L2.double().T).numpy(), cov.numpy(), atol=1e-5)
    with pytest.raises(Exception):
        synth.cholesky_L(-torch.eye(4))


def test_artefact_files_round_trip_in_the_reference_layout(tmp_path):
    """The reference's on-disk artefacts: ./bluenoise/cov_gaussian{BN,RN}_L_res64_d3.npz with key 'x'
    (iadb_bn.py:83-86) and <root>/noise/noise_batch{bs}_idx{:0>5}.npz with key 'noise' (iadb_bn.py:764,
    ddim_diffusers.py:667-669): files written by the helpers load through the reference's own expressions and back."""
    from bndm_b200 import io
    from bndm_b200.synth import hashed_tril, save_L_npz
    L = hashed_tril(seed=3)
    blue = tmp_path / "bluenoise"
    blue.mkdir()
    save_L_npz(str(blue / "cov_gaussianBN_L_res64_d3.npz"), L)
    save_L_npz(str(blue / "cov_gaussianRN_L_res64_d3.npz"), (L * np.float32(0.5)).astype(np.float32))
    # the reference's loader, verbatim (iadb_bn.py:83-86)
    ref_bn = np.load(str(blue / "cov_gaussianBN_L_res64_d3.npz"))["x"].astype(np.float32)
    assert np.array_equal(ref_bn, L)
    got_bn = io.load_cov_mat_L("gaussianBN", root=str(blue), device="cpu")
    got_rn = io.load_cov_mat_L("gaussianRN", root=str(blue), device="cpu")
    assert got_bn.dtype == torch.float32 and torch.equal(got_bn, torch.from_numpy(L))
    assert torch.equal(got_rn, torch.from_numpy(L) * 0.5)
    assert torch.equal(io.load_cov_mat_L("GBN", root=str(blue), device="cpu"), got_bn)       # everything but RN reads the BN file
    save_L_npz(str(blue / "cov_gaussianBN_L_res64_d3.npz"), L[:100, :100])
    with pytest.raises(ValueError):
        io.load_cov_mat_L("gaussianBN", root=str(blue), device="cpu")

    root = tmp_path / "run"
    (root / "noise").mkdir(parents=True)
    x0 = torch.randn(5, 3, 8, 8, dtype=torch.float64)                                        # np.random.randn is float64 (iadb_bn.py:761)
    io.save_noise_batch(str(root), 5, 17, x0)
    assert io.noise_batch_path(str(root), 5, 17).endswith("noise/noise_batch5_idx00017.npz")
    # the reference's reader, verbatim (iadb_bn.py:764-766): np.load(...)['noise'] -> torch -> .float()
    ref = torch.from_numpy(np.load(str(root / "noise" / "noise_batch5_idx00017.npz"))["noise"]).float()
    got = io.load_noise_batch(str(root), 5, 17, device="cpu")
    assert got.dtype == torch.float32 and torch.equal(got, ref) and torch.equal(got, x0.float())


def test_synthetic_red_and_blue_factors_have_the_right_spectra():
    """gaussianRN gets its own synthetic factor (low-pass), not the blue matrix: L L^T has unit diagonal and the power of
    L.z sits at low frequencies for red, at high frequencies for blue (the figure script's check, fig script :31-36)."""
    from bndm_b200.synth import blue_noise_L, red_noise_L
    rng = np.random.default_rng(0)
    z = rng.standard_normal((4096, 24))
    fy = np.fft.fftfreq(64)
    fr = np.sqrt(fy[:, None] ** 2 + fy[None, :] ** 2)
    ratios = {}
    for name, L in (("red", red_noise_L()), ("blue", blue_noise_L())):
        assert L.dtype == np.float32 and np.allclose(np.triu(L, 1), 0)
        assert np.allclose((L.astype(np.float64) ** 2).sum(1), 1.0, atol=1e-4)               # unit-variance pixels
        f = (L.astype(np.float64) @ z).T.reshape(24, 64, 64)
        spec = (np.abs(np.fft.fft2(f)) ** 2).mean(0)
        ratios[name] = spec[(fr > 0) & (fr < 0.1)].mean() / spec[fr > 0.35].mean()
    assert ratios["red"] > 20 and ratios["blue"] < 0.2, ratios
