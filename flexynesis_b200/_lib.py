"""ctypes binding of libfxn_b200.so (the C ABI declared in include/flexynesis_b200.h).

There is no fallback: if the shared library is missing the import of the engine fails with instructions, and
every wrapper raises on a non-zero return code with the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfxn_b200.so")

c_f32p = C.c_void_p
c_ll = C.c_longlong


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("lda", c_ll), ("a_mn_major", C.c_int),
        ("b_hi", C.c_void_p), ("b_lo", C.c_void_p), ("ldb", c_ll), ("b_mn_major", C.c_int),
        ("nterms", C.c_int),
        ("C", C.c_void_p), ("ldc", c_ll),
        ("bias", C.c_void_p),
        ("c_hi", C.c_void_p), ("c_lo", C.c_void_p), ("ldp", c_ll),
        ("colstats", C.c_void_p), ("stats_mode", C.c_int),
        ("splitk", C.c_int), ("block_n", C.c_int),
        ("epi_act", C.c_int), ("accumulate", C.c_int),
        ("alpha", C.c_float), ("alpha_dev", C.c_void_p),
        ("mse_x", C.c_void_p), ("ldx", c_ll), ("mse_acc", C.c_void_p),
        ("gauss_ra", C.c_void_p), ("gauss_rb", C.c_void_p), ("gauss_inv", C.c_float),
        ("stats_alpha", C.c_float), ("stats_alpha_dev", C.c_void_p),
        ("outputs_prezeroed", C.c_int),
        ("max_groups", C.c_int),
        ("fix_ws", C.c_void_p), ("fix_ws_bytes", c_ll),
        ("fix_flags", C.c_void_p), ("fix_flags_count", c_ll),
        ("rs_world", C.c_int), ("rs_rank", C.c_int), ("rs_per", c_ll), ("rs_base", C.c_void_p), ("rs_inbox", C.c_void_p),
    ]


class ReduceScatterContext:
    """Data-parallel jobs (parallel.NvlsDataParallel): while one of these is installed as `_lib.RS`, every weight-gradient
    GEMM whose plain stream-K output covers a contiguous, 4-aligned range of the gradient arena sends the reductions that
    belong to another rank's slice straight into that rank's inbox arena (fxn_gemm_desc.rs_*), and records the range so that
    fxn_dp_reduce_sumsq knows where own + inbox replaces the pull through the switch."""

    def __init__(self, world, rank, per, base_ptr, numel, inbox_ptrs):
        self.world, self.rank, self.per, self.base, self.numel = int(world), int(rank), int(per), int(base_ptr), int(numel)
        self.inbox = (C.c_void_p * self.world)(*[int(p) for p in inbox_ptrs])
        self.ranges = set()
        self.enabled = True
        self.min_elems = 1 << 18

    def claim(self, C_ptr, M, N, ldc):
        """(element offset) if this output can take the fused path, else None"""
        if not self.enabled or ldc != N:
            return None
        off = (int(C_ptr) - self.base) // 4
        n = int(M) * int(N)
        if (int(C_ptr) - self.base) % 16 or off < 0 or off + n > self.numel or n % 4 or off + n >= 1 << 31:
            return None
        return off

    def merged_ranges(self):
        out = []
        for lo, hi in sorted(self.ranges):
            if out and lo <= out[-1][1]:
                out[-1][1] = max(out[-1][1], hi)
            else:
                out.append([lo, hi])
        return out


RS = None        # the installed ReduceScatterContext (one data-parallel engine per process)


class BnFwdDesc(C.Structure):
    _fields_ = [
        ("V", C.c_void_p), ("ldv", c_ll), ("rows", c_ll), ("cols", C.c_int),
        ("partials", C.c_void_p), ("ntiles", C.c_int), ("tile_rows", C.c_int), ("partials_ld", C.c_int),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
        ("momentum", C.c_float), ("eps", C.c_float),
        ("train", C.c_int), ("act", C.c_int), ("p_drop", C.c_float),
        ("mask", C.c_void_p), ("ldm", c_ll), ("seed", C.c_ulonglong), ("seed_dev", C.c_void_p),
        ("out", C.c_void_p), ("ldo", c_ll),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("ldp", c_ll),
        ("saved", C.c_void_p),
        ("stat_rows", c_ll),
        ("keep_bits", C.c_void_p),
    ]


class BnBwdDesc(C.Structure):
    _fields_ = [
        ("V", C.c_void_p), ("ldv", c_ll), ("dOut", C.c_void_p), ("ldg", c_ll), ("rows", c_ll), ("cols", C.c_int),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("saved", C.c_void_p),
        ("act", C.c_int), ("p_drop", C.c_float), ("mask", C.c_void_p), ("ldm", c_ll), ("seed", C.c_ulonglong),
        ("seed_dev", C.c_void_p),
        ("pre_act", C.c_int),
        ("sums", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dbias", C.c_void_p),
        ("dV", C.c_void_p), ("ldd", c_ll),
        ("dv_hi", C.c_void_p), ("dv_lo", C.c_void_p), ("ldp", c_ll),
        ("grad_scale", C.c_float), ("accumulate_affine", C.c_int),
        ("stat_rows", c_ll), ("prezeroed", C.c_int), ("phase", C.c_int),
        ("keep_bits", C.c_void_p),
    ]


HEADS_MAX_VARS = 8


class HeadsVar(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("C", C.c_int), ("slot", C.c_int),
        ("W1", C.c_void_p), ("b1", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
        ("Wout", C.c_void_p), ("bout", C.c_void_p), ("y", C.c_void_p), ("logits", C.c_void_p),
        ("mask", C.c_void_p), ("ldm", c_ll), ("seed", C.c_ulonglong), ("log_var", C.c_void_p),
        ("dWout", C.c_void_p), ("dbout", C.c_void_p), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p),
    ]


class HeadsDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("L", C.c_int), ("sh", C.c_int), ("nv", C.c_int),
        ("F", C.c_void_p), ("ldf", c_ll), ("Zh", C.c_void_p), ("ldz", c_ll), ("G", C.c_void_p), ("ldg", c_ll),
        ("partials", C.c_void_p), ("saved", C.c_void_p), ("sums", C.c_void_p), ("acc", C.c_void_p),
        ("train", C.c_int), ("p_drop", C.c_float), ("momentum", C.c_float), ("eps", C.c_float),
        ("seed_dev", C.c_void_p), ("backward", C.c_int),
        ("var", HeadsVar * HEADS_MAX_VARS),
        ("dz_hi", C.c_void_p), ("dz_lo", C.c_void_p), ("ldzp", c_ll),
        ("df_hi", C.c_void_p), ("df_lo", C.c_void_p), ("ldfp", c_ll),
        ("dF", C.c_void_p), ("lddf", c_ll), ("dbias", C.c_void_p), ("zero_dbias", C.c_int),
    ]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"flexynesis_b200: {LIB_PATH} is missing. Build it with `make` (or `python -c 'import __graft_entry__ as g; "
            "g.build()'`) from the repository root; there is no CPU or PyTorch fallback for the training path.")
    lib = C.CDLL(LIB_PATH)
    lib.fxn_last_error.restype = C.c_char_p
    lib.fxn_launch_count.restype = c_ll
    lib.fxn_version.restype = C.c_int
    lib.fxn_gemm_fix_ws_bytes.restype = c_ll
    return lib


lib = _load()

# every exported symbol that include/flexynesis_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "fxn_version", "fxn_last_error", "fxn_launch_count", "fxn_reset_launch_count", "fxn_split_planes", "fxn_gemm",
    "fxn_gemm_stat_tiles", "fxn_gemm_plan", "fxn_gemm_fix_ws_bytes", "fxn_gemm_fix_flag_words", "fxn_bn_act_fwd", "fxn_bn_act_bwd", "fxn_col_stats", "fxn_head_out_fwd", "fxn_head_out_bwd",
    "fxn_cox_fwd", "fxn_cox_max_rows", "fxn_cox_fwd_ws", "fxn_cox_ws_floats", "fxn_heads_fused_ok", "fxn_heads_fwd", "fxn_heads_bwd", "fxn_total_loss", "fxn_triplet_fwd", "fxn_triplet_bwd", "fxn_clip_adam_step",
    "fxn_split_planes_multi", "fxn_gather_rows", "fxn_reparam_fwd", "fxn_reparam_bwd", "fxn_row_sqnorm",
    "fxn_mmd_finish", "fxn_mmd_grad", "fxn_loss_weights", "fxn_randn", "fxn_gcn_fwd", "fxn_gcn_bwd",
    "fxn_merge_col_stats", "fxn_merge_col_stats_big", "fxn_graph_gather", "fxn_graph_gather_ok", "fxn_node_lin_fwd", "fxn_node_lin_bwd", "fxn_debug_gemm_trace", "fxn_debug_gemm_cta_times", "fxn_dp_reduce_sumsq", "fxn_dp_adam_bcast", "fxn_dp_barrier",
]


class FxnError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise FxnError(f"{what or 'libfxn_b200'} failed ({rc}): {lib.fxn_last_error().decode()}")


def ptr(t) -> int:
    """Device pointer of a tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib.fxn_launch_count())


def reset_launch_count() -> None:
    lib.fxn_reset_launch_count()


def pad8(n: int) -> int:
    return (int(n) + 7) // 8 * 8


class Planes:
    """bf16 (hi, lo) operand planes of an fp32 matrix [rows x cols], shared leading dimension ld (elements)."""
    __slots__ = ("hi", "lo", "rows", "cols", "ld", "off")

    def __init__(self, hi, lo, rows, cols, ld, off=0):
        self.hi, self.lo, self.rows, self.cols, self.ld, self.off = hi, lo, rows, cols, ld, off

    @staticmethod
    def empty(rows, cols, device, ld=None):
        ld = ld or pad8(cols)
        hi = torch.zeros(rows * ld, dtype=torch.bfloat16, device=device)
        lo = torch.zeros(rows * ld, dtype=torch.bfloat16, device=device)
        return Planes(hi, lo, rows, cols, ld)

    def cols_view(self, c0, ncols):
        """Column window [c0, c0 + ncols) sharing storage (c0 % 8 == 0)."""
        assert c0 % 8 == 0
        return Planes(self.hi, self.lo, self.rows, ncols, self.ld, self.off + c0)

    def rows_view(self, r0, nrows):
        return Planes(self.hi, self.lo, nrows, self.cols, self.ld, self.off + r0 * self.ld)

    @property
    def hi_ptr(self):
        return self.hi.data_ptr() + 2 * self.off

    @property
    def lo_ptr(self):
        return self.lo.data_ptr() + 2 * self.off

    def to_float(self):
        h = self.hi[self.off:].float()
        l = self.lo[self.off:].float()
        full = (h + l)[: (self.rows - 1) * self.ld + self.cols]
        out = torch.empty(self.rows, self.cols, device=self.hi.device)
        for r in range(self.rows):   # debugging helper only
            out[r] = full[r * self.ld: r * self.ld + self.cols]
        return out


def fptr(t, off=0):
    """Pointer to element `off` of a float32 tensor."""
    return None if t is None else t.data_ptr() + 4 * off


def split_planes(src: torch.Tensor, dst: Planes) -> None:
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1
    check(lib.fxn_split_planes(C.c_void_p(src.data_ptr()), c_ll(src.stride(0)), c_ll(src.shape[0]), c_ll(src.shape[1]),
                               C.c_void_p(dst.hi_ptr), C.c_void_p(dst.lo_ptr), c_ll(dst.ld), C.c_void_p(stream())),
          "fxn_split_planes")


def gemm(M, N, K, a: Planes, a_mn, b: Planes, b_mn, *, C_ptr=None, ldc=0, bias=None, out: Planes = None,
         colstats=None, stats_mode=0, splitk=0, nterms=3, block_n=0, epi_act=0, accumulate=False, alpha=0.0,
         alpha_dev=None, mse_x=None, ldx=0, mse_acc=None, gauss_ra=None, gauss_rb=None, gauss_inv=0.0,
         stats_alpha=0.0, stats_alpha_dev=None, prezeroed=False, max_groups=0, fix: "FixWorkspace" = None) -> None:
    d = GemmDesc()
    d.M, d.N, d.K = int(M), int(N), int(K)
    d.a_hi, d.a_lo, d.lda, d.a_mn_major = a.hi_ptr, a.lo_ptr, a.ld, int(a_mn)
    d.b_hi, d.b_lo, d.ldb, d.b_mn_major = b.hi_ptr, b.lo_ptr, b.ld, int(b_mn)
    d.nterms = nterms
    d.C, d.ldc = C_ptr, int(ldc)
    d.bias = bias
    if out is not None:
        d.c_hi, d.c_lo, d.ldp = out.hi_ptr, out.lo_ptr, out.ld
    d.colstats, d.stats_mode = colstats, stats_mode
    d.splitk, d.block_n = splitk, block_n
    d.epi_act, d.accumulate = epi_act, int(accumulate)
    d.alpha, d.alpha_dev = alpha, alpha_dev
    d.mse_x, d.ldx, d.mse_acc = mse_x, int(ldx), mse_acc
    d.gauss_ra, d.gauss_rb, d.gauss_inv = gauss_ra, gauss_rb, gauss_inv
    d.stats_alpha, d.stats_alpha_dev = stats_alpha, stats_alpha_dev
    d.outputs_prezeroed = int(prezeroed)
    d.max_groups = int(max_groups)
    if fix is not None:
        d.fix_ws, d.fix_ws_bytes = fix.ws.data_ptr(), fix.ws.numel() * 4
        d.fix_flags, d.fix_flags_count = fix.flags.data_ptr(), fix.flags.numel()
    if (RS is not None and C_ptr and splitk < 0 and out is None and colstats is None and not epi_act and not accumulate
            and mse_x is None and bias is None):
        off = RS.claim(C_ptr, M, N, ldc)
        # only launches the planner runs as stream-K reductions anyway (the large weight gradients): forcing that mode on the
        # small ones cost more than their exchange saves
        if off is not None and int(M) * int(N) >= RS.min_elems and \
                gemm_plan(M, N, K, nterms, b_mn, plain_c=True, block_n=block_n, max_groups=max_groups)["streamk"] == 1:
            d.rs_world, d.rs_rank, d.rs_per, d.rs_base = RS.world, RS.rank, RS.per, RS.base
            d.rs_inbox = C.cast(RS.inbox, C.c_void_p)
            RS.ranges.add((off, off + int(M) * int(N)))
    check(lib.fxn_gemm(C.byref(d), C.c_void_p(stream())), "fxn_gemm")


class FixWorkspace:
    """Workspace of the GEMM's stream-K-with-fix-up mode (partial accumulators of split tiles + flags). Zeroed once; the
    kernel hands it back zeroed. One instance per chain of launches that cannot overlap in time (one per stream / graph
    branch)."""

    def __init__(self, device):
        self.ws = torch.zeros(int(lib.fxn_gemm_fix_ws_bytes()) // 4, dtype=torch.float32, device=device)
        self.flags = torch.zeros(int(lib.fxn_gemm_fix_flag_words()), dtype=torch.int32, device=device)


def gemm_plan(M, N, K, nterms=3, b_mn=0, plain_c=False, block_n=0, fix=False, max_groups=0) -> dict:
    """The launch plan fxn_gemm would choose (host-side cost model; callable without a GPU). plain_c: plain fp32 output
    (stream-K eligible); fix: a fix-up workspace is available (stream-K with fix-up eligible for fused epilogues)."""
    out = (C.c_int * 8)()
    mode = 1 if plain_c else (2 if fix else 0)
    check(lib.fxn_gemm_plan(C.c_int(M), C.c_int(N), C.c_int(K), C.c_int(nterms), C.c_int(int(b_mn)), C.c_int(mode),
                            C.c_int(int(block_n) | (int(max_groups) << 16)), out), "fxn_gemm_plan")
    keys = ("cta_group", "block_n", "stages", "streamk", "groups", "tiles_m", "tiles_n", "smem_bytes")
    return dict(zip(keys, (int(v) for v in out)))


def stat_tiles(M: int) -> int:
    return int(lib.fxn_gemm_stat_tiles(int(M)))


def col_stats(V_ptr, ldv, rows, cols, tile_rows, partials_ptr) -> None:
    check(lib.fxn_col_stats(C.c_void_p(V_ptr), c_ll(ldv), c_ll(rows), C.c_int(cols), C.c_int(tile_rows),
                            C.c_void_p(partials_ptr), C.c_void_p(stream())), "fxn_col_stats")


def bn_fwd(**kw) -> None:
    d = BnFwdDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    check(lib.fxn_bn_act_fwd(C.byref(d), C.c_void_p(stream())), "fxn_bn_act_fwd")


def bn_bwd(**kw) -> None:
    d = BnBwdDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    check(lib.fxn_bn_act_bwd(C.byref(d), C.c_void_p(stream())), "fxn_bn_act_bwd")


def head_out_fwd(D, ldd, rows, sh, W, bias, Cc, logits, ldl, kind, y, acc) -> None:
    check(lib.fxn_head_out_fwd(C.c_void_p(D), c_ll(ldd), C.c_int(rows), C.c_int(sh), C.c_void_p(W), C.c_void_p(bias),
                               C.c_int(Cc), C.c_void_p(logits), c_ll(ldl), C.c_int(kind), C.c_void_p(y),
                               C.c_void_p(acc), C.c_void_p(stream())), "fxn_head_out_fwd")


def head_out_bwd(D, ldd, rows, sh, W, Cc, logits, ldl, kind, y, acc, coef, weight, dD, ldg, dW, dbias,
                 prezeroed: bool = False) -> None:
    check(lib.fxn_head_out_bwd(C.c_void_p(D), c_ll(ldd), C.c_int(rows), C.c_int(sh), C.c_void_p(W), C.c_int(Cc),
                               C.c_void_p(logits), c_ll(ldl), C.c_int(kind), C.c_void_p(y), C.c_void_p(acc),
                               C.c_void_p(coef), C.c_void_p(weight), C.c_void_p(dD), c_ll(ldg), C.c_void_p(dW),
                               C.c_void_p(dbias), C.c_int(int(prezeroed)), C.c_void_p(stream())), "fxn_head_out_bwd")


def heads_fused_ok(Lt: int, sh: int, nv: int, max_c: int) -> bool:
    return bool(lib.fxn_heads_fused_ok(C.c_int(Lt), C.c_int(sh), C.c_int(nv), C.c_int(max_c)))


def heads_fwd(desc: HeadsDesc) -> None:
    check(lib.fxn_heads_fwd(C.byref(desc), C.c_void_p(stream())), "fxn_heads_fwd")


def heads_bwd(desc: HeadsDesc) -> None:
    check(lib.fxn_heads_bwd(C.byref(desc), C.c_void_p(stream())), "fxn_heads_bwd")


def cox_fwd(o, ldo, durations, events, n, coef, acc) -> None:
    check(lib.fxn_cox_fwd(C.c_void_p(o), c_ll(ldo), C.c_void_p(durations), C.c_void_p(events), C.c_int(n),
                          C.c_void_p(coef), C.c_void_p(acc), C.c_void_p(stream())), "fxn_cox_fwd")


def cox_fwd_ws(o, ldo, durations, events, n, coef, acc, workspace) -> None:
    """whole-chip pairwise version, no row limit; workspace: cox_ws_floats(n) fp32, 8-byte aligned"""
    check(lib.fxn_cox_fwd_ws(C.c_void_p(o), c_ll(ldo), C.c_void_p(durations), C.c_void_p(events), C.c_int(n),
                             C.c_void_p(coef), C.c_void_p(acc), C.c_void_p(workspace), C.c_void_p(stream())), "fxn_cox_fwd_ws")


def cox_ws_floats(n: int) -> int:
    lib.fxn_cox_ws_floats.restype = c_ll
    return int(lib.fxn_cox_ws_floats(C.c_int(n)))


def cox_max_rows() -> int:
    return int(lib.fxn_cox_max_rows())


def total_loss(n, acc, kinds, log_vars, dlog_vars, weighting, out) -> None:
    check(lib.fxn_total_loss(C.c_int(n), C.c_void_p(acc), C.c_void_p(kinds), C.c_void_p(log_vars),
                             C.c_void_p(dlog_vars), C.c_int(int(weighting)), C.c_void_p(out), C.c_void_p(stream())),
          "fxn_total_loss")


def triplet_fwd(A, P, N, ld, rows, L, margin, rowloss, acc) -> None:
    check(lib.fxn_triplet_fwd(C.c_void_p(A), C.c_void_p(P), C.c_void_p(N), c_ll(ld), C.c_int(rows), C.c_int(L),
                              C.c_float(margin), C.c_void_p(rowloss), C.c_void_p(acc), C.c_void_p(stream())),
          "fxn_triplet_fwd")


def triplet_bwd(A, P, N, ld, rows, L, rowloss, weight, dA, dP, dN, ldg, accumulate_a) -> None:
    check(lib.fxn_triplet_bwd(C.c_void_p(A), C.c_void_p(P), C.c_void_p(N), c_ll(ld), C.c_int(rows), C.c_int(L),
                              C.c_void_p(rowloss), C.c_void_p(weight), C.c_void_p(dA), C.c_void_p(dP), C.c_void_p(dN),
                              c_ll(ldg), C.c_int(int(accumulate_a)), C.c_void_p(stream())), "fxn_triplet_bwd")


def clip_adam(params, grads, m, v, n, lr, max_norm, grad_scale, sumsq, step, norm_out, beta1=0.9, beta2=0.999,
              eps=1e-8) -> None:
    check(lib.fxn_clip_adam_step(C.c_void_p(params), C.c_void_p(grads), C.c_void_p(m), C.c_void_p(v), c_ll(n),
                                 C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps), C.c_float(max_norm),
                                 C.c_float(grad_scale), C.c_void_p(sumsq), C.c_void_p(step), C.c_void_p(norm_out),
                                 C.c_void_p(stream())), "fxn_clip_adam_step")


def split_planes_multi(src, segs, nseg, max_elems, hi, lo) -> None:
    check(lib.fxn_split_planes_multi(C.c_void_p(src), C.c_void_p(segs), C.c_int(nseg), c_ll(max_elems), C.c_void_p(hi),
                                     C.c_void_p(lo), C.c_void_p(stream())), "fxn_split_planes_multi")


def gather_rows(src, ld_src, idx, nrows, cols, out, ldo, planes: "Planes" = None) -> None:
    check(lib.fxn_gather_rows(C.c_void_p(src), c_ll(ld_src), C.c_void_p(idx), c_ll(nrows), c_ll(cols), C.c_void_p(out),
                              c_ll(ldo), C.c_void_p(planes.hi_ptr if planes else None),
                              C.c_void_p(planes.lo_ptr if planes else None), c_ll(planes.ld if planes else 0),
                              C.c_void_p(stream())), "fxn_gather_rows")


def reparam_fwd(mean, s, eps, ld, rows, cols, z, zp: "Planes") -> None:
    check(lib.fxn_reparam_fwd(C.c_void_p(mean), C.c_void_p(s), C.c_void_p(eps), c_ll(ld), c_ll(rows), C.c_int(cols),
                              C.c_void_p(z), C.c_void_p(zp.hi_ptr), C.c_void_p(zp.lo_ptr), c_ll(zp.ld),
                              C.c_void_p(stream())), "fxn_reparam_fwd")


def reparam_bwd(dz, eps, ld, rows, cols, dm: "Planes", dsp: "Planes", dbias_mean, dbias_s) -> None:
    check(lib.fxn_reparam_bwd(C.c_void_p(dz), C.c_void_p(eps), c_ll(ld), c_ll(rows), C.c_int(cols),
                              C.c_void_p(dm.hi_ptr), C.c_void_p(dm.lo_ptr), C.c_void_p(dsp.hi_ptr),
                              C.c_void_p(dsp.lo_ptr), c_ll(dm.ld), C.c_void_p(dbias_mean), C.c_void_p(dbias_s),
                              C.c_void_p(stream())), "fxn_reparam_bwd")


def row_sqnorm(X, ld, rows, cols, out) -> None:
    check(lib.fxn_row_sqnorm(C.c_void_p(X), c_ll(ld), c_ll(rows), C.c_int(cols), C.c_void_p(out), C.c_void_p(stream())),
          "fxn_row_sqnorm")


def mmd_finish(cs_zz, cs_tt, cs_tz, mse_acc, dims, nlayers, B, P, acc) -> None:
    check(lib.fxn_mmd_finish(C.c_void_p(cs_zz), C.c_void_p(cs_tt), C.c_void_p(cs_tz), C.c_void_p(mse_acc),
                             C.c_void_p(dims), C.c_int(nlayers), C.c_int(B), C.c_int(P), C.c_void_p(acc),
                             C.c_void_p(stream())), "fxn_mmd_finish")


def mmd_grad(z, ldz, cs_zz, KZ, cs_tz, KT, ldk, nlayers, B, Lt, P, weight, dz, ldd) -> None:
    check(lib.fxn_mmd_grad(C.c_void_p(z), c_ll(ldz), C.c_void_p(cs_zz), C.c_void_p(KZ), C.c_void_p(cs_tz),
                           C.c_void_p(KT), c_ll(ldk), C.c_int(nlayers), C.c_int(B), C.c_int(Lt), C.c_int(P),
                           C.c_void_p(weight), C.c_void_p(dz), c_ll(ldd), C.c_void_p(stream())), "fxn_mmd_grad")


def loss_weights(n, log_vars, weighting, wts) -> None:
    check(lib.fxn_loss_weights(C.c_int(n), C.c_void_p(log_vars), C.c_int(int(weighting)), C.c_void_p(wts),
                               C.c_void_p(stream())), "fxn_loss_weights")


def randn(out, ld, rows, cols, seed, seed_dev=None) -> None:
    check(lib.fxn_randn(C.c_void_p(out), c_ll(ld), c_ll(rows), C.c_int(cols), C.c_ulonglong(seed), C.c_void_p(seed_dev),
                        C.c_void_p(stream())), "fxn_randn")


def gcn_fwd(X, B, N, Fin, rowptr, col, w, W, bias, emb, O, partials) -> None:
    check(lib.fxn_gcn_fwd(C.c_void_p(X), C.c_int(B), C.c_int(N), C.c_int(Fin), C.c_void_p(rowptr), C.c_void_p(col),
                          C.c_void_p(w), C.c_void_p(W), C.c_void_p(bias), C.c_int(emb), C.c_void_p(O),
                          C.c_void_p(partials), C.c_void_p(stream())), "fxn_gcn_fwd")


def gcn_bwd(X, dO, B, N, Fin, emb, csr_in, csr_out, W, dW, dbias, dX) -> None:
    """csr_in / csr_out: (rowptr, col, w) device int32/int32/fp32 tensors (by destination / by source)."""
    check(lib.fxn_gcn_bwd(C.c_void_p(X), C.c_void_p(dO), C.c_int(B), C.c_int(N), C.c_int(Fin), C.c_int(emb),
                          C.c_void_p(csr_in[0].data_ptr()), C.c_void_p(csr_in[1].data_ptr()),
                          C.c_void_p(csr_in[2].data_ptr()), C.c_void_p(csr_out[0].data_ptr()),
                          C.c_void_p(csr_out[1].data_ptr()), C.c_void_p(csr_out[2].data_ptr()), C.c_void_p(W),
                          C.c_void_p(dW), C.c_void_p(dbias), C.c_void_p(dX), C.c_void_p(stream())), "fxn_gcn_bwd")


def node_lin_fwd(X, B, N, Fin, Wr, emb, O, partials) -> None:
    check(lib.fxn_node_lin_fwd(C.c_void_p(X), C.c_int(B), C.c_int(N), C.c_int(Fin), C.c_void_p(Wr), C.c_int(emb),
                               C.c_void_p(O), C.c_void_p(partials), C.c_void_p(stream())), "fxn_node_lin_fwd")


def node_lin_bwd(X, dO, B, N, Fin, emb, Wr, dWr, dX) -> None:
    check(lib.fxn_node_lin_bwd(C.c_void_p(X), C.c_void_p(dO), C.c_int(B), C.c_int(N), C.c_int(Fin), C.c_int(emb),
                               C.c_void_p(Wr), C.c_void_p(dWr), C.c_void_p(dX), C.c_void_p(stream())), "fxn_node_lin_bwd")


def merge_col_stats(partials, ntiles, tile_rows, rows, cols, pld, merged) -> None:
    check(lib.fxn_merge_col_stats(C.c_void_p(partials), C.c_int(ntiles), C.c_int(tile_rows), c_ll(rows), C.c_int(cols),
                                  C.c_int(pld), C.c_void_p(merged), C.c_void_p(stream())), "fxn_merge_col_stats")


def dp_reduce_sumsq(mc_grad, grad_local, begin, end, scale, mc_partials, rank, scratch16, step, sync=None, peers=None,
                    inbox=None, ranges=None) -> None:
    """sync: None (the caller places the barrier) or (mc_flags, local_flags, epoch48, slot, world): barrier inside the kernel.
    peers: None (in-switch reduction through mc_grad) or the list of every rank's gradient-arena pointer (peer loads).
    inbox + ranges: this rank's inbox arena and the <= 8 element ranges [lo, hi) whose reduce-scatter already happened inside
    the weight-gradient GEMMs (ReduceScatterContext): there the sum is grad_local + inbox, and the inbox is cleared."""
    mc, lf, ep, slot, world = sync if sync is not None else (None, None, None, 0, 1)
    parr = (C.c_void_p * len(peers))(*peers) if peers else None
    nr = len(ranges) if (ranges and inbox) else 0
    rarr = (c_ll * (2 * nr))(*[int(v) for r in ranges for v in r]) if nr else None
    check(lib.fxn_dp_reduce_sumsq(C.c_void_p(mc_grad), C.c_void_p(grad_local), c_ll(begin), c_ll(end), C.c_float(scale),
                                  C.c_void_p(mc_partials), C.c_int(rank), C.c_void_p(scratch16), C.c_void_p(step),
                                  C.c_void_p(mc), C.c_void_p(lf), C.c_void_p(ep), C.c_int(slot), C.c_int(world),
                                  parr, C.c_int(len(peers) if peers else 0), C.c_void_p(inbox if nr else None), rarr,
                                  C.c_int(nr), C.c_void_p(stream())), "fxn_dp_reduce_sumsq")


def dp_adam_bcast(mc_param, param_local, grad_local, m, v, begin, end, partials, world, lr, max_norm, step, norm_out,
                  beta1=0.9, beta2=0.999, eps=1e-8, sync=None) -> None:
    mc, lf, ep, slot, _ = sync if sync is not None else (None, None, None, 0, 1)
    check(lib.fxn_dp_adam_bcast(C.c_void_p(mc_param), C.c_void_p(param_local), C.c_void_p(grad_local), C.c_void_p(m),
                                C.c_void_p(v), c_ll(begin), c_ll(end), C.c_void_p(partials), C.c_int(world), C.c_float(lr),
                                C.c_float(beta1), C.c_float(beta2), C.c_float(eps), C.c_float(max_norm), C.c_void_p(step),
                                C.c_void_p(norm_out), C.c_void_p(mc), C.c_void_p(lf), C.c_void_p(ep), C.c_int(slot),
                                C.c_void_p(stream())), "fxn_dp_adam_bcast")


def dp_barrier(mc_flags, local_flags, epoch, slot, world) -> None:
    check(lib.fxn_dp_barrier(C.c_void_p(mc_flags), C.c_void_p(local_flags), C.c_void_p(epoch), C.c_int(slot), C.c_int(world),
                             C.c_void_p(stream())), "fxn_dp_barrier")


def graph_gather_ok(N: int, Cc: int) -> bool:
    return bool(lib.fxn_graph_gather_ok(C.c_int(N), C.c_int(Cc)))


def graph_gather(inp, B, N, Cc, csr, out=None, out_planes: "Planes" = None) -> None:
    """csr: (rowptr, col, w[, order]) device int32 / int32 / fp32 / int32 tensors (order: nodes by decreasing row length)."""
    order = csr[3].data_ptr() if len(csr) > 3 and csr[3] is not None else None
    check(lib.fxn_graph_gather(C.c_void_p(inp), C.c_int(B), C.c_int(N), C.c_int(Cc), C.c_void_p(csr[0].data_ptr()),
                               C.c_void_p(csr[1].data_ptr()), C.c_void_p(csr[2].data_ptr()), C.c_void_p(order), C.c_void_p(out),
                               C.c_void_p(out_planes.hi_ptr if out_planes else None),
                               C.c_void_p(out_planes.lo_ptr if out_planes else None), C.c_void_p(stream())), "fxn_graph_gather")


def merge_col_stats_big(partials, ntiles, tile_rows, rows, cols, pld, fold, merged, scratch) -> None:
    check(lib.fxn_merge_col_stats_big(C.c_void_p(partials), C.c_int(ntiles), C.c_int(tile_rows), c_ll(rows), C.c_int(cols),
                                      C.c_int(pld), C.c_int(fold), C.c_void_p(merged), C.c_void_p(scratch),
                                      C.c_void_p(stream())), "fxn_merge_col_stats_big")
