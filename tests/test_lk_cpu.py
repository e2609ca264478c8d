"""The Python kernel DSL (triton_dist.lk) without a GPU: generated C++ text, nvcc cross-compilation for sm_100a (incl. the tcgen05 GEMM
ladder: the SASS must contain the Blackwell tensor-core / TMA instructions), and the CPU interpreter against PyTorch.
Reference test strategy: python/little_kernel/tests/unit/* (pure-python codegen tests that run without a GPU)."""
import os
import shutil
import subprocess

import pytest
import torch

from triton_dist import lk
from triton_dist.lk import CompileError, ll
from triton_dist.lk.kernels import simt as K

N_CONST = 4


def _body(kernel) -> str:
    src = kernel.cuda_source()
    return src[src.index('extern "C" __global__'):src.index('extern "C" int lk_launch')]


def test_codegen_types_and_scoping():
    @lk.kernel(block=64)
    def k(x: ll.ptr[ll.bf16], y: ll.ptr[ll.f32], n: ll.i32, s: ll.u32):
        i = ll.blockIdx.x * 64 + ll.threadIdx.x
        if i < n:
            v = x[i] * 2            # bf16 * literal -> computed in fp32
        else:
            v = 0.0                 # same variable in both arms: ONE declaration at function scope (Python scoping)
        w = s >> 1                  # unsigned stays unsigned, the literal adopts the type
        q = i // 3 + i % 3          # integer floor-div / mod
        r = i / 4                   # true division of ints -> float
        y[i] = v + w + q + r
    body = _body(k)
    assert "__launch_bounds__(64)" in body and "__nv_bfloat16* x" in body
    assert body.count("float v;") == 1 and "uint32_t w;" in body and "int q;" in body and "float r;" in body
    assert "((float)(x[i])) * 2.0f" in body
    assert "(s >> 1u)" in body and "(i / 3)" in body and "(i % 3)" in body and "((float)(i) / (float)(4))" in body


def test_codegen_constants_static_loops_and_pruning():
    FLAG = False

    @lk.kernel
    def k(y: ll.ptr[ll.i32]):
        acc: ll.i32 = 0
        for j in ll.static_range(N_CONST):          # unrolled by the code generator, j is a constant in the body
            acc += j * 10
        if FLAG:                                    # closure constant: the branch is pruned
            acc = -1
        if N_CONST > 2 and ll.threadIdx.x == 0:     # constant part of the condition folds away
            y[0] = acc + (N_CONST << 2)
    body = _body(k)
    assert "for (" not in body and "acc = (acc + 30);" in body and "-1" not in body
    assert "if ((((int)threadIdx.x) == 0))" in body and "(acc + 16)" in body


def test_codegen_device_function_specialisation():
    def scale(v, f: ll.constexpr):
        return v * f

    @lk.kernel
    def k(a: ll.ptr[ll.f32], b: ll.ptr[ll.i32]):
        a[0] = scale(a[1], 2.0)
        a[2] = scale(a[3], 2.0)          # same types + same constant: one specialisation
        b[0] = scale(b[1], 3)            # int operand, other constant: a second one
    src = k.cuda_source()
    assert src.count("__device__ __forceinline__ float lk_scale__0(float v)") == 1
    assert src.count("__device__ __forceinline__ int lk_scale__1(int v)") == 1
    assert src.count("lk_scale__0(") == 3 and "(v * 3)" in src and "(v * 2.0f)" in src


def test_codegen_shared_memory_layout_and_intrinsics():
    @lk.kernel(block=128, cluster=(2, 1, 1))
    def k(t: ll.TmaDescriptor, out: ll.ptr[ll.u32]):
        ll.align_memory(1024)
        tile = ll.dyn_shared([2, 128 * 64], ll.bf16, align=1024)
        bar = ll.dyn_shared([2], ll.u64)
        st = ll.shared([4, 8], ll.f32)
        st[1, 2] = 1.0
        if ll.warp_id() == 0 and ll.elect_one():
            ll.mbar_init(bar + 1, 1)
            ll.fence_barrier_init()
            ll.mbar_arrive_expect_tx(bar + 1, 128 * 64 * 2)
            ll.tma_load_2d(t, bar + 1, tile[1], 0, 0)
        ll.cluster_sync()
        ll.mbar_wait(bar + 1, 0)
        out[0] = ll.smem_addr(tile[1]) + ll.cluster_rank()
    body = _body(k)
    assert "const __grid_constant__ CUtensorMap t" in body
    assert "reinterpret_cast<__nv_bfloat16*>(lk_dyn + 0)" in body and "reinterpret_cast<uint64_t*>(lk_dyn + 32768)" in body
    assert "__shared__ __align__(16) float st[32];" in body and "st[10] = 1.0f;" in body
    assert "td::ptx::tma_load_2d(&t, (bar + (1)), (tile + 8192), 0, 0)" in body
    assert k.dyn_smem_bytes == 32768 + 16 + 1024          # carve + alignment slack
    assert "lk_dyn_raw + ((1024u - (td::ptx::smem_u32(lk_dyn_raw) & 1023u)) & 1023u)" in body


def test_compile_errors_are_located():
    @lk.kernel
    def undefined(y: ll.ptr[ll.f32]):
        y[0] = nope        # noqa: F821

    @lk.kernel
    def bad_break(y: ll.ptr[ll.f32]):
        for j in ll.static_range(4):
            if y[j] > 0:
                break

    @lk.kernel
    def runtime_constexpr(y: ll.ptr[ll.i32]):
        n: ll.constexpr = y[0]

    @lk.kernel
    def no_annotation(y):
        pass

    @lk.kernel
    def two_types(y: ll.ptr[ll.u32]):
        for v in range(4):
            y[v] = 0
        v = ll.make_uint4(1, 2, 3, 4)          # noqa: F841  (a local has ONE type per function)

    for kern, frag in ((undefined, "'nope' is not defined"), (bad_break, "statically unrolled"), (runtime_constexpr, "constexpr"),
                       (no_annotation, "type annotation"), (two_types, "use another name")):
        with pytest.raises(CompileError) as e:
            kern.cuda_source()
        assert frag in str(e.value) and kern.name in str(e.value)


def test_inline_asm_and_casts():
    @lk.kernel
    def k(p: ll.ptr[ll.u32], q: ll.ptr[ll.f32]):
        v: ll.u32 = 0
        ll.asm("mov.u32 %0, %%smid;", outputs=[v], memory=False)
        ll.asm("red.release.sys.global.add.u32 [%0], %1;", inputs=[p, v])
        q[0] = ll.uint_as_float(ll.u32(ll.i64(v) << 3))
        b = ll.ptr_cast(q, ll.u8)
        b[1] = 7
    body = _body(k)
    assert 'asm volatile("mov.u32 %0, %%smid;" : "=r"(v) : )' in body
    assert '"l"(p), "r"(v) : "memory")' in body
    assert "reinterpret_cast<uint8_t*>(q)" in body and "((uint32_t)((((int64_t)(v)) << 3ll)))" in body


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(shutil.which("nvcc") is None and not __import__("os").path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_nvcc_compiles_examples_and_gemm_ladder_emits_tcgen05():
    for kern in (K.saxpy, K.block_sum, K.softmax_rows, K.histogram, K.ring_shift, K.allgather_push):
        kern.compile()
    from triton_dist import jit
    from triton_dist.lk.kernels.gemm_sm100 import get_gemm
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    for cg, needles in ((1, ("UTCHMMA", "UTMALDG.2D", "LDTM.x32", "UTCBAR")),
                        (2, ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST", "UTCATOMSWS.2CTA"))):
        g = get_gemm(256, 4, cg)
        g.compile()
        so = g._lib._name
        sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
        assert "sm_100a" in sass
        for n in needles:
            assert n in sass, (cg, n)
    from triton_dist.lk.kernels.ag_gemm import make_ag_gemm
    ag = make_ag_gemm(256, 4, 4)                               # comm CTAs + tcgen05 tiles in one kernel
    ag.compile()
    sass = subprocess.run([cuobjdump, "-sass", ag._lib._name], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG.2D" in sass and "td::notify" in ag.cuda_source() and "td::wait_ge<true>" in ag.cuda_source()
    from triton_dist.lk.kernels.gemm_rs import make_gemm_rs
    rs = make_gemm_rs(256, 4, 4)                               # tcgen05 tiles + bf16x2 reductions into the owner + collector CTAs
    rs.compile()
    sass = subprocess.run([cuobjdump, "-sass", rs._lib._name], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "REDG.E.ADD.BF16" in sass and "REDG.E.ADD.STRONG.SYS" in sass        # tile reductions + counter release-add
    from triton_dist.lk.kernels.gemm_sm100 import make_gemm_persistent
    pk = make_gemm_persistent(256, 6, 2)                       # the persistent rung: two TMEM accumulators, ring across tiles
    pk.compile()
    sass = subprocess.run([cuobjdump, "-sass", pk._lib._name], capture_output=True, text=True).stdout
    for n in ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST", "LDTM.x32"):
        assert n in sass, n
    assert pk.dyn_smem_bytes < 227 * 1024
    with pytest.raises(AssertionError):
        make_gemm_persistent(256, 6, 1)                        # 48 KB per stage x 6 does not fit one SM
    assert jit._CACHE.exists()


# ------------------------------------------------------------------------------------------------------------
def test_interpreter_matches_torch():
    torch.manual_seed(0)
    x, y = torch.randn(1000), torch.randn(1000)
    ref = 2.5 * x + y
    K.saxpy.interpret((1000 + 127) // 128, x, y, 2.5, 1000)
    torch.testing.assert_close(y, ref)

    xb = torch.randn(3000).bfloat16()
    out = torch.zeros(1)
    K.block_sum.interpret(4, xb, out, 3000)                 # shuffles + shared memory + atomics, 4 blocks x 128 threads
    torch.testing.assert_close(out[0], xb.float().sum(), atol=1e-2, rtol=1e-4)

    m = torch.randn(5, 300)
    sm = torch.empty_like(m)
    K.softmax_rows.interpret(5, m, sm, 300)
    torch.testing.assert_close(sm, m.softmax(-1), atol=1e-6, rtol=1e-5)

    ids = torch.randint(0, 60, (2000,), dtype=torch.int32)
    cnt = torch.zeros(60, dtype=torch.int32)
    K.histogram.interpret(3, ids, cnt, 2000, 60)            # dynamic shared memory
    assert torch.equal(cnt, torch.bincount(ids.long(), minlength=60).int())


def test_interpreter_integer_wraparound_and_errors():
    assert ll.u32(-1) == 0xFFFFFFFF and ll.i32(0x80000000) == -(1 << 31) and ll.u8(257) == 1 and ll.bf16(1.001) == 1.0

    @lk.kernel(block=32)
    def boom(y: ll.ptr[ll.f32]):
        if ll.threadIdx.x == 3:
            ll.trap()
        ll.syncthreads()

    with pytest.raises(RuntimeError, match="trap"):
        boom.interpret(1, torch.zeros(1))                    # one thread fails: the block's barrier is aborted, no hang
    with pytest.raises(NotImplementedError):
        ll.ld_shared_v4(None)                                # raw shared-window addresses have no CPU meaning


def test_gdn_chunk_kernels_in_the_interpreter_match_the_recurrence():
    """The chunked gated-delta-rule forward written in the DSL (prepare + scan kernels, WY form with an in-kernel triangular solve):
    executed by the CPU interpreter it reproduces the token-by-token recurrence, including a non-zero initial state and a ragged tail."""
    from triton_dist.lk.kernels.gdn_chunk import chunk_gated_delta_rule_lk, get_kernels
    from triton_dist.ops.gdn import gated_delta_rule_recurrent
    torch.manual_seed(0)
    B, T, H, DK, DV, C = 1, 40, 2, 16, 16, 16
    q = torch.randn(B, T, H, DK) * 0.5
    k = torch.nn.functional.normalize(torch.randn(B, T, H, DK), dim=-1)
    v = torch.randn(B, T, H, DV) * 0.5
    g, beta = -torch.rand(B, T, H) * 0.5, torch.rand(B, T, H)
    s0 = torch.randn(B, H, DK, DV) * 0.1
    o, S = chunk_gated_delta_rule_lk(q, k, v, g, beta, initial_state=s0, chunk_size=C, interpret=True)
    ro, rS = gated_delta_rule_recurrent(q, k, v, g, beta, initial_state=s0)
    torch.testing.assert_close(o, ro, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(S, rS, atol=1e-5, rtol=1e-5)
    # the production shape cross-compiles: 64-token chunks, 128 x 128 heads, bf16 I/O, ~100 KB of dynamic shared memory
    prep, scan = get_kernels(64, 128, 128, 32, ll.bf16)
    prep.compile(); scan.compile()
    assert 90_000 < prep.dyn_smem_bytes < 110_000 and scan.dyn_smem_bytes < 30_000


def _clamp(v, lo, hi):
    return min(max(v, lo), hi)


def _lerp(a, b, t: ll.constexpr):
    return a + (b - a) * t


def test_statement_coverage_codegen_and_interpreter_agree_with_python():
    """while / continue / break, tuple swap, negative and unrolled ranges, conditional expressions, 64-bit pointer offsets, bit
    operators, struct fields: the generated C++ compiles with nvcc and the interpreter gives the values plain Python arithmetic gives."""
    @lk.kernel(block=64)
    def k1(x: ll.ptr[ll.f32], y: ll.ptr[ll.i32], z: ll.ptr[ll.u8], n: ll.i32, big: ll.i64, flag: ll.bool_):
        tid = ll.threadIdx.x
        i = tid
        acc: ll.f32 = 0.0
        cnt = 0
        while i < n:
            if x[i] < 0.0:
                i += 64
                continue
            acc += x[i]
            cnt += 1
            if cnt > 1000:
                break
            i += 64
        a, b = cnt, tid
        a, b = b, a
        y[tid] = a * N_CONST + b
        for j in range(10, 0, -2):
            acc -= 0.5
        for j in ll.unroll(range(4)):
            acc += _lerp(1.0, 3.0, 0.5)
        z[tid] = ll.u8(_clamp(cnt, 0, 255)) if flag else ll.u8(0)
        x[tid] = acc if not (tid == 0 or tid == 63) else -acc
        p = x + big
        if p != x and tid == 0:
            y[64] = ll.i32(big >> 3) & 0xFF
        y[65 + tid] = (tid << 2) ^ (tid % 3) | (1 if tid > 5 else 0)
        v4 = ll.make_uint4(1, 2, 3, 4)
        v4.x = ll.u32(tid)
        y[200 + tid] = ll.i32(v4.x + v4.w)

    body = _body(k1)
    for frag in ("while ((i < n)) {", "continue;", "break;", "for (j = 10; j > 0; j += -2) {", "#pragma unroll", "acc = (acc + 2.0f);",
                 "lk_t3 = b;", "(flag ? ((uint8_t)(lk__clamp__0(cnt, 0, 255))) : ((uint8_t)0))", "p = (x + (big));", "(big >> 3ll)",
                 "v4.x = ((uint32_t)(tid));", "int64_t big, bool flag"):
        assert frag in body, frag
    k1.compile()
    x = torch.randn(256)
    xr = x.clone()
    y = torch.zeros(512, dtype=torch.int32)
    z = torch.zeros(64, dtype=torch.uint8)
    k1.interpret(1, x, y, z, 256, 16, True)
    cnt = torch.stack([(xr[t::64] >= 0).sum() for t in range(64)])
    t = torch.arange(64)
    assert torch.equal(y[:64], (t * N_CONST + cnt).int()) and torch.equal(z, cnt.clamp(0, 255).to(torch.uint8))
    assert int(y[64]) == 2 and torch.equal(y[65:129], ((t << 2) ^ (t % 3) | (t > 5).long()).int()) and torch.equal(y[200:264], (t + 4).int())
    acc = torch.stack([xr[tt::64][xr[tt::64] >= 0].sum() for tt in range(64)]) - 2.5 + 8.0
    acc[0], acc[63] = -acc[0], -acc[63]
    torch.testing.assert_close(x[:64], acc, atol=1e-5, rtol=1e-5)


def test_gemm_ladder_runs_in_the_cpu_pipeline_model():
    """The tcgen05 GEMM ladder executed by the interpreter's functional model of the Blackwell pipeline (mbarrier phases and tx counts,
    TMA, tensor memory, cta_group::1 / ::2 MMAs and multicast commits): every rung -- including the persistent one with two accumulators
    and a ring that runs across tiles -- reproduces A @ B^T on ragged shapes, i.e. no wrong phase, stage or tile index, no deadlock."""
    from triton_dist.lk.kernels.gemm_sm100 import run_gemm, run_gemm_persistent
    torch.manual_seed(0)

    def check(fn, M, N, K, **kw):
        a, b = (torch.randn(M, K) * 0.5).bfloat16(), (torch.randn(N, K) * 0.5).bfloat16()
        c = fn(a, b, **kw)
        torch.testing.assert_close(c.float(), a.float() @ b.float().t(), atol=0.15, rtol=2e-2)

    check(run_gemm, 200, 264, 320, cta_group=1, BN=128, STAGES=2)       # ragged M / N, 5 k-blocks through a 2-stage ring
    check(run_gemm, 300, 392, 192, cta_group=2, STAGES=2)               # 2-CTA pairs, 2 x 2 tiles
    check(run_gemm_persistent, 520, 328, 192, num_sms=2)                # ONE cluster walks 3 x 2 tiles: ring + accumulators wrap
    # the ladder as a table: a single-stage level, a 1-CTA persistent level and the endpoint, on one ragged problem
    from triton_dist.lk.kernels.gemm_sm100 import LEVELS, test_all_levels
    assert sorted(LEVELS) == list(range(1, 10))
    res = test_all_levels(200, 264, 192, device="cpu", levels=(1, 6, 9), num_sms=2)
    assert set(res) == {1, 6, 9} and all(err < 0.15 for err, _ in res.values()), res


def test_pipeline_model_reports_protocol_errors(monkeypatch):
    """A barrier that can never complete is reported as a deadlock with the barrier's state; a warp reading TMEM lanes it does not own
    and an MMA issued by the wrong CTA are rejected."""
    from triton_dist.lk import pipeline as P
    monkeypatch.setattr(P, "TIMEOUT_S", 0.5)

    @lk.kernel(block=32)
    def stuck(y: ll.ptr[ll.f32]):
        bar = ll.dyn_shared([1], ll.u64)
        if ll.threadIdx.x == 0:
            ll.mbar_init(bar, 2)            # two arrivals expected ...
            ll.mbar_arrive(bar)             # ... one ever comes
        ll.syncthreads()
        ll.mbar_wait(bar, 0)

    with pytest.raises(P.Deadlock, match="pending arrivals 1"):
        stuck.interpret(1, torch.zeros(1))

    @lk.kernel(block=64)
    def wrong_lanes(y: ll.ptr[ll.f32]):
        regs = ll.local([32], ll.u32)
        ll.tmem_ld_32x32b_x32(ll.u32(0), regs)      # warp 1 must read lanes 32..63

    with pytest.raises(RuntimeError, match="may only read TMEM lanes 32"):
        wrong_lanes.interpret(1, torch.zeros(1))


def test_debug_wait_variant_guards_spin_loops(monkeypatch):
    """TD_DEBUG_WAITS=<ms>: kernels compiled against the device headers get timer-guarded waits that print a diagnostic and trap
    (hang detection); without it the hooks vanish."""
    from triton_dist import _build
    src = r"""
#include "td/primitives.cuh"
using namespace td;
__global__ void waits(SymmCtx c, uint32_t* flags, uint32_t* slots) {
  wait<true, true>(flags, 4, 7u);
  if (threadIdx.x == 0) wait_ge<true>(flags + 8, 3u);
  barrier_all_block(c, slots, 5u);
}
extern "C" void launch_waits(SymmCtx c, void* f, void* s, void* stream) { waits<<<1, 64, 0, (cudaStream_t)stream>>>(c, (uint32_t*)f, (uint32_t*)s); }
"""
    from triton_dist import jit
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    monkeypatch.delenv("TD_DEBUG_WAITS", raising=False)
    assert _build.debug_wait_timeout_ns() is None
    plain = subprocess.run([cuobjdump, "-sass", jit.compile_cuda(src, name="waits")._name], capture_output=True, text=True).stdout
    monkeypatch.setenv("TD_DEBUG_WAITS", "2500")
    assert _build.debug_wait_timeout_ns() == 2_500_000_000
    guarded = subprocess.run([cuobjdump, "-sass", jit.compile_cuda(src, name="waits")._name], capture_output=True, text=True).stdout
    assert "GLOBALTIMER" not in plain and "BPT.TRAP" not in plain
    assert "GLOBALTIMER" in guarded and "BPT.TRAP" in guarded


def test_aot_export_of_dsl_kernels(tmp_path):
    """tools/compile_aot: a DSL kernel becomes <name>.cu + <name>.so + a C header whose prototype a plain C compiler accepts."""
    from triton_dist.tools import compile_aot
    compile_aot.main(["triton_dist.lk.kernels.simt:block_sum", "triton_dist.lk.kernels.gemm_sm100:get_gemm(256, 4, 2)", "--out", str(tmp_path)])
    for name in ("block_sum", "lk_gemm_bn256_s4_cg2"):
        assert (tmp_path / f"{name}.so").stat().st_size > 10_000 and (tmp_path / f"{name}.cu").exists()
    hdr = (tmp_path / "block_sum.h").read_text()
    assert "int lk_launch_block_sum(void* x /* __nv_bfloat16* */, float* out, int n, unsigned gx" in hdr
    assert "const void* tA /* CUtensorMap on the host */" in (tmp_path / "lk_gemm_bn256_s4_cg2.h").read_text()
    (tmp_path / "use.c").write_text('#include "block_sum.h"\nint f(void* x, float* o, int n, void* s) { return lk_launch_block_sum(x, o, n, 4, 1, 1, 128, 1, 1, 1, 1, 1, 0, s); }\n')
    r = subprocess.run(["gcc", "-c", str(tmp_path / "use.c"), "-I", str(tmp_path), "-o", str(tmp_path / "use.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import ctypes
    assert hasattr(ctypes.CDLL(str(tmp_path / "block_sum.so")), "lk_launch_block_sum")


PRIMES = (2, 3, 5, 7, 11, 13, 17, 19)
WEIGHTS = [0.5, 0.25, 0.125, 0.0625]


def test_python_sequences_become_lookup_tables():
    """A module-level list / tuple of numbers indexed with a run-time value is emitted once as a constant table (compile-time indices
    still fold); the interpreter indexes the Python object."""
    @lk.kernel(block=32)
    def k(y: ll.ptr[ll.f32], z: ll.ptr[ll.i32]):
        t = ll.threadIdx.x
        z[t] = PRIMES[t % 8] + PRIMES[3]
        y[t] = WEIGHTS[t & 3] * 2.0 + WEIGHTS[t % 4]
    body = _body(k)
    assert body.count("const int lk_tbl0[8] = {2, 3, 5, 7, 11, 13, 17, 19};") == 1 and "const float lk_tbl1[4] = {0.5f, 0.25f, 0.125f, 0.0625f};" in body
    assert "(lk_tbl0[(t % 8)] + 7)" in body and body.count("lk_tbl1[") == 3            # one declaration, two uses
    k.compile()
    y, z = torch.zeros(32), torch.zeros(32, dtype=torch.int32)
    k.interpret(1, y, z)
    t = torch.arange(32)
    assert torch.equal(z, (torch.tensor(PRIMES)[t % 8] + 7).int())
    torch.testing.assert_close(y, torch.tensor(WEIGHTS)[t % 4] * 3.0)


def test_warp_mma_linear_matches_matmul_in_the_interpreter():
    """The megakernel's tensor-core LINEAR algorithm as a DSL kernel (mma.sync m16n8k16 with the interpreter's model of the PTX
    fragment layout): fragment-ordered staging, the shared K permutation of weights and activations, K split across warps with a
    shared-memory reduction, several column passes, ragged batches -- against x @ W^T."""
    from triton_dist.lk.kernels.linear_mma import make_linear_mma, run_linear_mma
    torch.manual_seed(0)
    for (B, K, N, tn, KC) in ((24, 96, 32, 16, 64),          # 2 groups -> 4 warps share one group along K; chunks of 64 + 32
                              (40, 64, 64, 64, 64),           # 8 groups: one per warp
                              (17, 128, 128, 128, 128)):      # 16 groups: two passes
        x, W = (torch.randn(B, K) * 0.5).bfloat16(), (torch.randn(N, K) * 0.5).bfloat16()
        o = run_linear_mma(x, W, tile_n=tn, KC=KC, interpret=True)
        torch.testing.assert_close(o.float(), x.float() @ W.float().t(), atol=6e-2, rtol=2e-2)
    k = make_linear_mma(256)
    k.compile()
    sass = subprocess.run([shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump", "-sass", k._lib._name], capture_output=True, text=True).stdout
    assert sass.count("HMMA") >= 8


def _attn_ref(q, k, v, causal, scale, softcap):
    B, S, Hq, _ = q.shape
    G = Hq // k.shape[2]
    qf, kf, vf = q.float().permute(0, 2, 1, 3), k.float().permute(0, 2, 1, 3).repeat_interleave(G, 1), v.float().permute(0, 2, 1, 3).repeat_interleave(G, 1)
    s = qf @ kf.transpose(-1, -2) * scale
    if softcap > 0:
        s = softcap * torch.tanh(s / softcap)
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), 1), float("-inf"))
    return (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3)


def test_flash_attention_on_mma_sync_in_the_interpreter():
    """The megakernel's FLASH_ATTN (prefill) algorithm as a DSL kernel: A fragments straight from 16-byte query loads with the shared
    d permutation, K row-major / V transposed-as-key-pairs in shared memory, scores kept in registers between the two MMAs, online
    softmax over quads -- against the fp32 reference.  Covers GQA, a packed qkv tensor (strided views), two KV tiles with a ragged tail,
    causal and full masks, the soft cap, and the reinterpreting ``ll.ptr_cast`` the epilogue stores through."""
    from triton_dist.lk.kernels.flash_mma import make_flash_mma, run_flash_mma
    torch.manual_seed(0)
    B, S, Hq, Hkv = 1, 66, 1, 1
    qkv = torch.randn(B, S, Hq + 2 * Hkv, 128).bfloat16()
    q, k, v = qkv[:, :, :Hq], qkv[:, :, Hq:Hq + Hkv], qkv[:, :, Hq + Hkv:]
    o = run_flash_mma(q, k, v, causal=True, threads=32, interpret=True)
    torch.testing.assert_close(o.float(), _attn_ref(q, k, v, True, 128 ** -0.5, 0.0), atol=2e-2, rtol=2e-2)
    B, S, Hq, Hkv = 1, 20, 2, 1                     # GQA: two query heads share one KV head
    q, k, v = (torch.randn(B, S, Hq, 128) * 2).bfloat16(), torch.randn(B, S, Hkv, 128).bfloat16(), torch.randn(B, S, Hkv, 128).bfloat16()
    o = run_flash_mma(q, k, v, causal=False, softcap=5.0, sm_scale=0.2, threads=32, interpret=True)
    torch.testing.assert_close(o.float(), _attn_ref(q, k, v, False, 0.2, 5.0), atol=2e-2, rtol=2e-2)
    kern = make_flash_mma(256)
    kern.compile()
    sass = subprocess.run([shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump", "-sass", kern._lib._name], capture_output=True, text=True).stdout
    assert sass.count("HMMA") >= 128 and "STL" not in sass          # 64 + 64 MMAs per KV tile, no spills


def test_language_extra_and_stdlib_in_dsl_kernels(tmp_path):
    """The reference's helper vocabulary as DSL intrinsics (scoped ld / st / atomics / red / fences lower to ONE PTX instruction each),
    the device-side binary searches, the grid barrier and extern_call: generated CUDA compiles for sm_100a, the interpreter gives the
    same answers; the AOT registry exports a registered kernel."""
    from triton_dist.lk import language_extra as le, stdlib

    @lk.kernel(block=64)
    def le_selftest(cnt: ll.ptr[ll.u32], flag: ll.ptr[ll.u32], out: ll.ptr[ll.u32], big: ll.ptr[ll.u64], fl: ll.ptr[ll.f32],
                    sorted_vals: ll.ptr[ll.i32], n_sorted: ll.i32, bar: ll.ptr[ll.u32]):
        t = le.tid(0)
        n = le.ntid(0)
        old = le.atomic_add(cnt, 1, scope="gpu", semantic="relaxed")
        out[ll.blockIdx.x * 80 + t] = old
        le.red_release(cnt + 1, 2, scope="sys")
        le.__syncthreads()
        if t == 0 and ll.blockIdx.x == 0:
            le.st(flag, 7, scope="sys", semantic="release")
            le.fence("acq_rel", "sys")
            le.membar("gpu")
            big[0] = le.pack_b32_v2(3, 4)
            out[n] = le.atomic_cas(cnt + 2, 0, 99, scope="sys", semantic="acq_rel")
            out[n + 3] = le.unpack_hi(big[0])
            le.atomic_add(fl, 1.5)
        if t == 1 and ll.blockIdx.x == 0:
            le.wait_eq(flag, 7)
            out[n + 1] = le.ld(flag, scope="sys", semantic="acquire") + le.ld_acquire(flag, "gpu")
        v = le.__shfl_down_sync_i32(0xFFFFFFFF, t, 1)
        b = le.__ballot_sync(0xFFFFFFFF, t % 2 == 0)
        w = le.atomic_add_per_warp(cnt + 3, 1)
        if t == 5 and ll.blockIdx.x == 0:
            out[n + 2] = ll.u32(v) + (b & 0xF) + w * 0 + le.laneid()
            out[n + 4] = stdlib.bisect_left(sorted_vals, n_sorted, 7) * 100 + stdlib.bisect_right(sorted_vals, n_sorted, 7)
            out[n + 5] = ll.u32(stdlib.extern_call("td::ptx::bf16_hi", ll.f32, ll.u32(0x40400000)))       # bf16 3.0 in the high half
        stdlib.grid_barrier(bar, ll.gridDim.x)
        if t == 0:
            out[ll.blockIdx.x * 80 + 79] = le.ld(cnt, scope="gpu", semantic="acquire")        # every block sees every block's adds

    src = le_selftest.cuda_source()
    for needle in ("atom.relaxed.gpu.global.add.u32", "red.release.sys.global.add.u32", "st.release.sys.global.b32", "fence.acq_rel.sys",
                   "membar.gl", "atom.acq_rel.sys.global.cas.b32", "ld.acquire.sys.global.b32", "atom.relaxed.gpu.global.add.f32",
                   "td::grid_barrier", "td::ptx::bf16_hi("):
        assert needle in src, needle
    le_selftest.compile()
    cnt, flag, out = torch.zeros(8, dtype=torch.int32), torch.zeros(1, dtype=torch.int32), torch.zeros(160, dtype=torch.int32)
    big, fl, bar = torch.zeros(1, dtype=torch.int64), torch.zeros(1), torch.zeros(1, dtype=torch.int32)
    vals = torch.tensor([1, 3, 7, 7, 7, 9, 12], dtype=torch.int32)
    stdlib.EXTERN_INTERP["td::ptx::bf16_hi"] = lambda w: torch.tensor([int(w) >> 16], dtype=torch.int16).view(torch.bfloat16).float().item()
    le_selftest.interpret(2, cnt, flag, out, big, fl, vals, vals.numel(), bar)
    assert cnt[:4].tolist() == [128, 256, 99, 4] and flag.item() == 7 and fl.item() == 1.5 and big.item() == (4 << 32) | 3
    assert sorted(out[:64].tolist() + out[80:144].tolist()) == list(range(128))
    assert out[64:70].tolist() == [0, 14, 16, 4, 2 * 100 + 5, 3] and out[79].item() == 128 and out[159].item() == 128

    from triton_dist.tools import compile_aot as A

    @A.aot_compile_spaces({"saxpy_aot": {}})
    @lk.kernel(block=128)
    def saxpy_for_aot(x: ll.ptr[ll.f32], y: ll.ptr[ll.f32], a: ll.f32, n: ll.i32):
        i = ll.blockIdx.x * 128 + ll.threadIdx.x
        if i < n:
            y[i] = a * x[i] + y[i]

    infos = A.export_registered(tmp_path, names=["saxpy_aot"])
    assert len(infos) == 1 and os.path.exists(infos[0]["so"]) and "lk_launch_saxpy_for_aot" in open(infos[0]["header"]).read()
    A.AOT_REGISTRY.pop("saxpy_aot")


def test_dsl_microbenchmarks_compile_and_check_out_in_the_interpreter():
    """The micro-benchmark suite written in the DSL (triton_dist/lk/bench: FMA / SFU / mma.sync throughput, FMA and pointer-chase
    latencies, sync latency, global copy / read, shared-memory bank conflicts, shuffles, integer IPC, occupancy probe): every kernel
    cross-compiles for sm_100a and, at toy sizes in the interpreter, produces the closed-form output its runner checks."""
    from triton_dist.lk.bench import BENCHES, KERNELS, run_all
    assert len(KERNELS) == 13 and set(BENCHES) == set(KERNELS)
    for k in KERNELS.values():
        k.compile()
    res = run_all(interpret=True)
    assert set(res) == set(BENCHES) and all(r["ok"] for r in res.values()), {n: r["ok"] for n, r in res.items()}
    sass = subprocess.run([shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump", "-sass", KERNELS["mma_sync_throughput"]._lib._name],
                          capture_output=True, text=True).stdout
    assert sass.count("HMMA") >= 4 and "MUFU.EX2" in subprocess.run(
        [shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump", "-sass", KERNELS["sfu_throughput"]._lib._name], capture_output=True, text=True).stdout


def test_decode_attention_kernels_in_the_interpreter():
    """Split-KV GQA decode + log-sum-exp combine written in the DSL (the megakernel's ATTN task / flash-decode algorithm): one split writes
    the output directly, several splits go through (m, l, o) partials; ragged lengths; the returned LSE feeds a cross-rank merge."""
    from triton_dist.lk.kernels.flash_decode import decode_combine, decode_split, gqa_decode_lk
    torch.manual_seed(0)
    B, Hq, Hkv = 1, 4, 2
    q = (torch.randn(B, Hq, 128) * 0.5).bfloat16()
    k, v = (torch.randn(B, 40, Hkv, 128) * 0.5).bfloat16(), (torch.randn(B, 40, Hkv, 128) * 0.5).bfloat16()
    lens = torch.tensor([29], dtype=torch.int32)
    ro, rl = torch.zeros(B, Hq, 128), torch.zeros(B, Hq)
    for h in range(Hq):
        s = (q[0, h].float() @ k[0, :29, h // 2].float().t()) * 128 ** -0.5
        ro[0, h], rl[0, h] = torch.softmax(s, -1) @ v[0, :29, h // 2].float(), torch.logsumexp(s, -1)
    o1 = gqa_decode_lk(q, k, v, lens, n_splits=1, interpret=True)
    o2, l2 = gqa_decode_lk(q, k, v, lens, n_splits=2, interpret=True, return_lse=True)
    torch.testing.assert_close(o1.float(), ro, atol=5e-3, rtol=5e-3)
    torch.testing.assert_close(o2.float(), ro, atol=5e-3, rtol=5e-3)
    torch.testing.assert_close(l2, rl, atol=1e-4, rtol=1e-4)
    decode_split.compile()
    decode_combine.compile()


def test_sass_fingerprints_of_the_dsl_communication_kernels():
    """What the Python communication kernels become on sm_100a: switch reductions (``LDGMC...ADD.BF16x8`` = multimem.ld_reduce), tcgen05 + TMA
    next to them in the fused GEMM + AllReduce, single 64-bit system-scope stores / loads for the flag-in-data atoms, GPU-scope atomics and
    reductions for the EP slot counters and the grid barrier."""
    from triton_dist.lk.kernels.all_to_all import all_to_all_ll
    from triton_dist.lk.kernels.allgather_ll import allgather_ll
    from triton_dist.lk.kernels.allreduce_nvls import make_allreduce_nvls
    from triton_dist.lk.kernels.ep_a2a import ep_dispatch
    from triton_dist.lk.kernels.gemm_ar import make_gemm_ar

    def sass(k):
        k.compile()
        return subprocess.run([shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump", "-sass", k._lib._name], capture_output=True, text=True).stdout

    one, two = make_allreduce_nvls(ll.bf16)
    for k in (one, two):
        s_ = sass(k)
        assert "LDGMC.E.HPADD.BF16x8" in s_ and "MEMBAR.ALL.SYS" in s_ and "STG.E.128" in s_
    s_ = sass(make_gemm_ar(128, 4, 8))
    assert "UTCHMMA" in s_ and "UTMALDG.2D" in s_ and "LDGMC.E.HPADD.BF16x8" in s_
    s_ = sass(allgather_ll)
    assert "STG.E.64.STRONG.SYS" in s_ and "LDG.E.64.STRONG.SYS" in s_ and "MEMBAR" not in s_          # no fence, no barrier: the atom is the flag
    assert "STG.E.64.STRONG.SYS" in sass(all_to_all_ll)
    s_ = sass(ep_dispatch)
    assert "ATOMG.E.ADD" in s_ and "REDG.E.ADD.STRONG.GPU" in s_

