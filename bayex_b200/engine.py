"""Device-side plumbing between the bayex-shaped Python API and the C ABI.

torch is used for device memory, streams and (optionally) ``torch.distributed``; all arithmetic happens in
``libbayex_b200.so``.  ``Posterior`` owns the cached factor that the reference recomputes on every
``predict`` call (gp.py:41-49).
"""
from __future__ import annotations

import ctypes as C
from contextlib import contextmanager

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

_STREAMS = {}
_WORKSPACES = {}


def device(index=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("bayex_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device() if index is None else index)


def stream(dev: torch.device) -> torch.cuda.Stream:
    """One private non-default stream per device (CUDA-graph capture is illegal on the legacy stream)."""
    s = _STREAMS.get(dev.index)
    if s is None:
        s = _STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return s


@contextmanager
def on_stream(st: torch.cuda.Stream, dev: torch.device):
    """Run on the library's private stream, ordered after the caller's current stream on entry and before it on exit,
    so tensors handed in or returned can be used on the caller's stream without further synchronisation."""
    cur = torch.cuda.current_stream(dev)
    st.wait_stream(cur)
    with torch.cuda.stream(st):
        yield
    cur.wait_stream(st)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _workspace(dev, nbytes) -> torch.Tensor:
    ws = _WORKSPACES.get(dev.index)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _WORKSPACES.pop(dev.index, None)
        ws = _WORKSPACES[dev.index] = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    return ws


def dims_array(domain: dict):
    """bayex.domain objects (dict order = column order, domain.py:175) -> bx_dim_t[d]."""
    arr = (_lib.DimT * len(domain))()
    for j, dom in enumerate(domain.values()):
        if getattr(dom, "is_integer", False):
            arr[j] = _lib.DimT(1, float(dom.lower), float(dom.upper), int(dom.lower), int(dom.upper))
        else:
            arr[j] = _lib.DimT(0, float(np.float32(dom.lower)), float(np.float32(dom.upper)), 0, 0)
    return arr


def raw_vector(params, d) -> np.ndarray:
    """GPParams (python floats or (1,1)/(1,d) arrays: gp.py:41 broadcasting, quirk Q2) -> float32[d+2]."""
    nr = np.asarray(params.noise, dtype=np.float32).reshape(-1)[0]
    ar = np.asarray(params.amplitude, dtype=np.float32).reshape(-1)[0]
    ls = np.asarray(params.lengthscale, dtype=np.float32).reshape(-1)
    if ls.size == 1 and d != 1:
        ls = np.full((d,), ls[0], dtype=np.float32)
    if ls.size != d:
        raise ValueError(f"lengthscale has {ls.size} entries for {d} input dimensions")
    return np.concatenate([[nr], [ar], ls]).astype(np.float32)


def compact(x, y, mask):
    """Apply the reference's mask on the host side: keep valid rows only (see include/bayex_b200.h)."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[:, None]  # tests/test_gp.py:20-31: 1-D inputs mean d = 1
    y = np.asarray(y, dtype=np.float32).reshape(-1)
    m = np.asarray(mask).astype(bool).reshape(-1)
    if not m.any():
        raise ValueError("no valid observation (mask is all False)")
    return np.ascontiguousarray(x[m]), np.ascontiguousarray(y[m])


class Posterior:
    """The fitted GP on one GPU: compacted observations, raw hypers and (after ``build``) the factor."""

    def __init__(self, x, y, raw, dev=None, want_tc=True):
        self.dev = device() if dev is None else dev
        self.stream = stream(self.dev)
        x = np.ascontiguousarray(x, dtype=np.float32)
        self.n, self.d = x.shape
        if self.d > _lib.MAX_D:
            raise ValueError(f"at most {_lib.MAX_D} dimensions are supported, got {self.d}")
        self.np_ = (self.n + _lib.TILE - 1) // _lib.TILE * _lib.TILE
        self.d4 = (self.d + 3) // 4 * 4
        self.want_tc = want_tc        # fp16 head/tail copies of T: the production tcgen05 engine
        with on_stream(self.stream, self.dev):
            self.X = torch.from_numpy(x).to(self.dev, non_blocking=False)
            self.y = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(self.dev)
            self.raw = torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float32)).to(self.dev)
        self.factor = None
        self._bufs = None

    # -- fit path ---------------------------------------------------------------------------------------------
    def _ws(self):
        nbytes = lib.bx_fit_workspace_bytes(self.n, self.d)
        return _workspace(self.dev, nbytes), nbytes

    def fit_hypers(self, lr=1e-3, steps=100, trace=False, use_graph=True):
        """gp.py:90-110 on the device; updates ``self.raw`` in place."""
        ws, nbytes = self._ws()
        with on_stream(self.stream, self.dev):
            tr = torch.empty((steps, self.d + 2), dtype=torch.float32, device=self.dev) if trace else None
            check(lib.bx_fit_adamw(_ptr(self.X), _ptr(self.y), self.n, self.d, _ptr(self.raw), lr, steps, _ptr(tr),
                                   _ptr(ws), nbytes, 1 if use_graph else 0, C.c_void_p(self.stream.cuda_stream)),
                  "bx_fit_adamw")
        self.factor = None
        return tr

    def fit_hypers_batch(self, raws, lr=1e-3, steps=100):
        """``posterior_fit`` (gp.py:90-110) from every row of ``raws`` (R, d + 2) side by side on this GPU (one CUDA
        graph with R parallel branches per step).  Returns ``(fitted raws (R, d + 2), rows (R, 2 + 2(d + 2)))`` as host
        arrays; ``rows[r]`` is the ``nlml_grad`` record at the fitted hypers of restart r.  ``self.raw`` is untouched."""
        raws = np.ascontiguousarray(raws, dtype=np.float32)
        R, P = raws.shape
        assert P == self.d + 2 and 1 <= R <= _lib.FIT_BATCH_MAX
        nbytes = lib.bx_fit_workspace_bytes(self.n, self.d) * R
        ws = _workspace(self.dev, nbytes)
        with on_stream(self.stream, self.dev):
            d_raws = torch.from_numpy(raws).to(self.dev)
            out = torch.empty((R, 2 + 2 * P), dtype=torch.float32, device=self.dev)
            check(lib.bx_fit_adamw_batch(_ptr(self.X), _ptr(self.y), self.n, self.d, _ptr(d_raws), R, lr, steps, _ptr(out),
                                         _ptr(ws), nbytes, C.c_void_p(self.stream.cuda_stream)), "bx_fit_adamw_batch")
            return d_raws.cpu().numpy(), out.cpu().numpy()

    def nlml_grad(self):
        """(nlml, loss, dloss/draw, dnlml/draw) at the current raw hypers (gp.py:74-87)."""
        ws, nbytes = self._ws()
        with on_stream(self.stream, self.dev):
            out = torch.empty(2 + 2 * (self.d + 2), dtype=torch.float32, device=self.dev)
            check(lib.bx_nlml_grad(_ptr(self.X), _ptr(self.y), self.n, self.d, _ptr(self.raw), _ptr(out), _ptr(ws),
                                   nbytes, C.c_void_p(self.stream.cuda_stream)), "bx_nlml_grad")
            o = out.cpu().numpy()
        P = self.d + 2
        return float(o[0]), float(o[1]), o[2:2 + P].copy(), o[2 + P:2 + 2 * P].copy()

    def build(self):
        """gp.py:41-49 once: factor arrays for every later scoring call."""
        ws, nbytes = self._ws()
        np_, d4 = self.np_, self.d4
        with on_stream(self.stream, self.dev):
            if self._bufs is None:
                self._bufs = self._alloc_bufs()
            b = self._bufs
            check(lib.bx_factor_build(_ptr(self.X), _ptr(self.y), self.n, self.d, _ptr(self.raw), _ptr(b["U"]),
                                      _ptr(b["T"]), _ptr(b["T16hi"]), _ptr(b["T16lo"]),
                                      _ptr(b["alpha"]), _ptr(b["scale"]), _ptr(b["hyp"]), _ptr(ws), nbytes,
                                      C.c_void_p(self.stream.cuda_stream)),
                  "bx_factor_build")
        self._set_factor()
        self._warn_if_not_pd()
        return self

    def _warn_if_not_pd(self):
        # gp.py:48: the reference's cholesky returns NaN silently on a non-PD Gram matrix; so do the kernels (hyp[7] records
        # it).  Surface it as a warning - never an exception, the NaNs propagate exactly as they do upstream.
        rc = lib.bx_check_status(_ptr(self._bufs["hyp"]), C.c_void_p(self.stream.cuda_stream))
        if rc == _lib.E_NOTPD:
            import warnings
            warnings.warn("bayex_b200: " + lib.bx_last_error().decode("utf-8", "replace") + "; posterior values are NaN",
                          RuntimeWarning, stacklevel=3)
        elif rc != 0:
            check(rc, "bx_check_status")

    def can_append(self) -> bool:
        """A built factor with a free padding row (np = roundup(n, 128)): ``append`` applies."""
        return self.factor is not None and self.n < self.np_

    def append(self, x_new, y_new):
        """One more observation at UNCHANGED hypers (optimizer.py:261-269): the factor grows by one
        bordered row in O(n^2) - two triangular matrix-vector products - instead of the O(n^3) rebuild of gp.py:46-49."""
        if not self.can_append():
            raise ValueError("append() needs a built factor with a free padding row; rebuild the posterior instead")
        x_new = np.ascontiguousarray(x_new, dtype=np.float32).reshape(-1)
        if x_new.size != self.d:
            raise ValueError(f"the new point has {x_new.size} coordinates for {self.d} input dimensions")
        nbytes = lib.bx_factor_append_workspace_bytes(self.np_)
        ws = _workspace(self.dev, nbytes)
        with on_stream(self.stream, self.dev):
            xn = torch.from_numpy(x_new).to(self.dev)
            self.X = torch.cat([self.X, xn[None]], dim=0)
            self.y = torch.cat([self.y, torch.tensor([float(np.float32(y_new))], dtype=torch.float32, device=self.dev)])
            check(lib.bx_factor_append(C.byref(self.factor), _ptr(xn), _ptr(self.y), _ptr(ws), nbytes,
                                       C.c_void_p(self.stream.cuda_stream)), "bx_factor_append")
        self.n += 1
        self._set_factor()
        self._warn_if_not_pd()
        return self

    def _alloc_bufs(self):
        """Everything a scoring rank needs lives in ONE flat allocation (segments 256-byte aligned), so the multi-GPU
        hand-out of a freshly built factor is a single NCCL broadcast: hyp | scale | alpha | raw | U | T | T16hi | T16lo."""
        np_, d4 = self.np_, self.d4
        seg = [("hyp", 16), ("scale", d4), ("alpha", np_), ("rawcopy", self.d + 2), ("U", np_ * d4), ("T", np_ * np_)]
        if self.want_tc:
            seg += [("T16hi", np_ * np_ // 2), ("T16lo", np_ * np_ // 2)]  # fp16 matrices, counted in float32 words
        off, offs = 0, {}
        for name, nfloat in seg:
            offs[name] = (off, nfloat)
            off += (nfloat + 63) // 64 * 64
        flat = torch.zeros(off, dtype=torch.float32, device=self.dev)
        view = lambda name: flat[offs[name][0]:offs[name][0] + offs[name][1]]
        b = dict(flat=flat, hyp=view("hyp"), scale=view("scale"), alpha=view("alpha"), rawcopy=view("rawcopy"),
                 U=view("U").view(np_, d4), T=view("T").view(np_, np_), T16hi=None, T16lo=None)
        if self.want_tc:
            b["T16hi"] = view("T16hi").view(torch.float16).view(np_, np_)
            b["T16lo"] = view("T16lo").view(torch.float16).view(np_, np_)
        return b

    def _set_factor(self):
        b = self._bufs
        p = lambda t: 0 if t is None else t.data_ptr()
        self.factor = _lib.FactorT(self.n, self.np_, self.d, self.d4, p(b["U"]), p(b["T"]),
                                   p(b["T16hi"]), p(b["T16lo"]), p(b["alpha"]), p(b["scale"]), p(b["hyp"]))

    def hyp(self) -> np.ndarray:
        self._need_factor()
        with on_stream(self.stream, self.dev):
            return self._bufs["hyp"].cpu().numpy()

    def raw_host(self) -> np.ndarray:
        with on_stream(self.stream, self.dev):
            return self.raw.cpu().numpy()

    def set_raw(self, raw):
        """Replace the raw hypers in place (same observations, new starting point - restarts of the fit)."""
        with on_stream(self.stream, self.dev):
            self.raw.copy_(torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float32)))
        self.factor = None

    def _need_factor(self):
        if self.factor is None:
            self.build()

    # -- multi-GPU: the factor is built on one rank and broadcast (NCCL over NVLink) --------------------------
    def broadcast(self, src=0, group=None):
        """One collective for the whole factor (flat buffer), stream-ordered on the library stream: no host sync."""
        import torch.distributed as dist
        with on_stream(self.stream, self.dev):
            if self._bufs is None:
                self._bufs = self._alloc_bufs()
            b = self._bufs
            b["rawcopy"].copy_(self.raw)
            dist.broadcast(b["flat"], src=src, group=group)
            self.raw.copy_(b["rawcopy"])
        self._set_factor()
        return self

    # -- scoring ----------------------------------------------------------------------------------------------
    def score_points(self, cand, acq="UCB", param=0.0, ystar=0.0, want=("mu", "sigma", "acq"), best=False):
        """Posterior moments / acquisition at explicit points (gp.py:59-71, acq.py).  Returns device tensors."""
        self._need_factor()
        with on_stream(self.stream, self.dev):
            if not isinstance(cand, torch.Tensor):
                cand = torch.from_numpy(np.ascontiguousarray(cand, dtype=np.float32)).to(self.dev)
            cand = cand.contiguous()
            M = cand.shape[0]
            out = {k: (torch.empty(M, dtype=torch.float32, device=self.dev) if k in want else None)
                   for k in ("mu", "sigma", "acq")}
            bst = torch.zeros(1, dtype=torch.int64, device=self.dev) if best else None
            # point sets of at least BX_POINTS_TC_MIN take the two-tier tcgen05 path: it needs the hand-over workspace
            ws, nbytes = None, 0
            if M >= _lib.POINTS_TC_MIN and self.want_tc:
                nbytes = lib.bx_score_workspace_bytes(M)
                ws = _workspace(self.dev, nbytes)
            check(lib.bx_score_points(C.byref(self.factor), _ptr(cand), M, _lib.ACQ_IDS[acq], float(param), float(ystar),
                                      _ptr(out["mu"]), _ptr(out["sigma"]), _ptr(out["acq"]), _ptr(bst), 0, _ptr(ws), nbytes,
                                      C.c_void_p(self.stream.cuda_stream)), "bx_score_points")
        out["best"] = bst
        return out

    def score_generated(self, key, dims, m_begin, m_count, acq, param, ystar, keep_acq=True, engine=_lib.ENGINE_AUTO,
                        best=None):
        """Fused Threefry -> posterior -> acquisition -> arg-max over a slice of the candidate draw."""
        self._need_factor()
        with on_stream(self.stream, self.dev):
            vals = torch.empty(m_count, dtype=torch.float32, device=self.dev) if keep_acq else None
            if best is None:
                best = torch.zeros(1, dtype=torch.int64, device=self.dev)
            nbytes = lib.bx_score_workspace_bytes(int(m_count))
            ws = _workspace(self.dev, nbytes)
            check(lib.bx_score_generated(C.byref(self.factor), _lib.key_words(key), dims, int(m_begin), int(m_count),
                                         _lib.ACQ_IDS[acq], float(param), float(ystar), _ptr(vals), _ptr(best),
                                         int(engine), _ptr(ws), nbytes, C.c_void_p(self.stream.cuda_stream)), "bx_score_generated")
        return vals, best

    # -- fused sample() pieces: everything below enqueues on the library stream and returns device tensors -------------
    def sample_front(self, key, dims, m_begin, m_count, acq, param, ystar, k, engine=_lib.ENGINE_AUTO, keep_acq=False):
        """optimizer.py:183-197 + :205 for one candidate shard: Threefry -> posterior -> acquisition -> per-CTA top-k ->
        merge.  Returns ``(list, vals, idx, acq)``: the (k + 1)-entry key list (see include/bayex_b200.h), its valid
        entries unpacked ascending at the front of ``vals`` / ``idx``, and the acquisition vector if ``keep_acq``."""
        self._need_factor()
        with on_stream(self.stream, self.dev):
            lst = torch.zeros(k + 1, dtype=torch.int64, device=self.dev)
            vals = torch.empty(k, dtype=torch.float32, device=self.dev)
            idx = torch.empty(k, dtype=torch.int32, device=self.dev)
            acq_out = torch.empty(m_count, dtype=torch.float32, device=self.dev) if keep_acq else None
            self.last_counts = torch.zeros(2, dtype=torch.int32, device=self.dev)  # [valid list entries, handed over]
            self.last_m_count = int(m_count)
            if m_count > 0:
                nbytes = lib.bx_score_topk_workspace_bytes(int(m_count), int(k))
                ws = _workspace(self.dev, nbytes)
                check(lib.bx_score_generated_topk(C.byref(self.factor), _lib.key_words(key), dims, int(m_begin), int(m_count),
                                                  _lib.ACQ_IDS[acq], float(param), float(ystar), int(k), _ptr(lst), _ptr(vals),
                                                  _ptr(idx), _ptr(self.last_counts), _ptr(acq_out), int(engine), _ptr(ws), nbytes,
                                                  C.c_void_p(self.stream.cuda_stream)), "bx_score_generated_topk")
        return lst, vals, idx, acq_out

    def merge_lists(self, lists: torch.Tensor, n_lists: int, k: int):
        """Merge the all-gathered key lists of the candidate shards into one; same outputs as above."""
        with on_stream(self.stream, self.dev):
            out = torch.empty(k + 1, dtype=torch.int64, device=self.dev)
            vals = torch.empty(k, dtype=torch.float32, device=self.dev)
            idx = torch.empty(k, dtype=torch.int32, device=self.dev)
            nbytes = 8 * n_lists * k
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            check(lib.bx_topk_merge(_ptr(lists), int(n_lists), int(k), _ptr(out), _ptr(vals), _ptr(idx), None, _ptr(ws), nbytes,
                                    C.c_void_p(self.stream.cuda_stream)), "bx_topk_merge")
        return out, vals, idx

    def propose(self, key, dims, opt_x, opt_val, S, best: torch.Tensor):
        """optimizer.py:203-207 on the device: arg-max over concat(refined, random), the winner's raw coordinates.
        Returns a device tensor [d + 2]: point, value, position in the concatenation (uint32 bits)."""
        with on_stream(self.stream, self.dev):
            out = torch.empty(self.d + 2, dtype=torch.float32, device=self.dev)
            check(lib.bx_propose(_lib.key_words(key), dims, self.d, _ptr(opt_x), _ptr(opt_val), int(S), _ptr(best), _ptr(out),
                                 C.c_void_p(self.stream.cuda_stream)), "bx_propose")
        return out

    def topk(self, vals: torch.Tensor, k: int, index_base=0):
        with on_stream(self.stream, self.dev):
            M = vals.numel()
            nbytes = lib.bx_topk_workspace_bytes(M, k)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            ov = torch.empty(k, dtype=torch.float32, device=self.dev)
            oi = torch.empty(k, dtype=torch.int32, device=self.dev)
            check(lib.bx_topk(_ptr(vals), M, k, int(index_base), _ptr(ov), _ptr(oi), _ptr(ws), nbytes,
                              C.c_void_p(self.stream.cuda_stream)), "bx_topk")
        return ov, oi

    def gather(self, key, dims, idx: torch.Tensor):
        with on_stream(self.stream, self.dev):
            k = idx.numel()
            out = torch.empty((k, self.d), dtype=torch.float32, device=self.dev)
            check(lib.bx_threefry_gather(_lib.key_words(key), dims, self.d, _ptr(idx), k, _ptr(out),
                                         C.c_void_p(self.stream.cuda_stream)), "bx_threefry_gather")
        return out

    def refine(self, starts: torch.Tensor, rounds, acq, param, ystar):
        self._need_factor()
        with on_stream(self.stream, self.dev):
            x = starts.clone().contiguous()
            S = x.shape[0]
            nbytes = lib.bx_refine_workspace_bytes(C.byref(self.factor), S)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            val = torch.empty(S, dtype=torch.float32, device=self.dev)
            check(lib.bx_refine(C.byref(self.factor), _ptr(x), S, int(rounds), _lib.ACQ_IDS[acq], float(param),
                                float(ystar), _ptr(val), _ptr(ws), nbytes, C.c_void_p(self.stream.cuda_stream)),
                  "bx_refine")
        return x, val


def threefry_fill(key, dims, d, m_begin, m_count, dev=None) -> torch.Tensor:
    """Candidates [m_begin, m_begin+m_count) of ``sample_params(key, (M,))`` as a device tensor [m_count, d]."""
    dev = device() if dev is None else dev
    st = stream(dev)
    with on_stream(st, dev):
        out = torch.empty((m_count, d), dtype=torch.float32, device=dev)
        check(lib.bx_threefry_fill(_lib.key_words(key), dims, d, int(m_begin), int(m_count), _ptr(out),
                                   C.c_void_p(st.cuda_stream)), "bx_threefry_fill")
    return out


def domain_sample(key, dims, count, dev=None) -> torch.Tensor:
    """``Domain.sample(key, (count,))``: one domain, the key used as is (domain.py:74,130)."""
    dev = device() if dev is None else dev
    st = stream(dev)
    with on_stream(st, dev):
        out = torch.empty((count,), dtype=torch.float32, device=dev)
        check(lib.bx_domain_sample(_lib.key_words(key), dims, 0, int(count), _ptr(out), C.c_void_p(st.cuda_stream)),
              "bx_domain_sample")
    return out


def to_unsigned(packed_i64: int) -> int:
    return packed_i64 & 0xFFFFFFFFFFFFFFFF
