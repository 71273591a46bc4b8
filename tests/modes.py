"""IntersectWithMode (tutorials/verify/rtcore_helpers.h:751-904): one ray array funnelled through every flavour of the
query API -- rtcIntersect1 / 1M / 1Mp / 4 / 8 / 16 / NM / Np and their rtcOccluded twins -- so that every entry point is
checked against the same expectation.  Works with any library exporting the rtc* ABI (the product and, to prove the
harness, the reference library).  With `device=True` the SoA / pointer layouts are placed in GPU memory (product only)."""
import ctypes as C

import numpy as np

import cases

rt = cases.rt
MODES = ("1", "1M", "1Mp", "4", "8", "16", "NM", "Np")
RAY_NAMES = rt.RAY_DTYPE.names
HIT_NAMES = tuple(n for n in rt.RAYHIT_DTYPE.names if n not in RAY_NAMES)


def _sigs(L):
    vp, u, sz = C.c_void_p, C.c_uint, C.c_size_t
    ctxp = C.POINTER(rt.IntersectContext)
    for name in ("rtcIntersectNp", "rtcOccludedNp"):
        f = getattr(L, name)
        f.restype, f.argtypes = None, [vp, ctxp, vp, u]


def _soa_block(rays, names, n_pad):
    """(len(names), n_pad) uint32 block, padding lanes inactive (tnear = +inf, tfar = -inf)."""
    n = len(rays)
    blk = np.zeros((len(names), n_pad), dtype=np.uint32)
    for k, nm in enumerate(names):
        blk[k, :n] = rays[nm].view(np.uint32)
        if nm == "tnear":
            blk[k, n:] = np.float32(np.inf).view(np.uint32)
        if nm == "tfar":
            blk[k, n:] = np.float32(-np.inf).view(np.uint32)
    return blk


def _unsoa(rays, blk, names):
    n = len(rays)
    for k, nm in enumerate(names):
        rays[nm] = blk[k, :n].view(rays[nm].dtype)


class _Mem:
    """Host (numpy, 64-byte aligned) or device (torch) storage for uint32 words with the same little interface."""

    def __init__(self, words, device):
        self.device = device
        words = np.ascontiguousarray(words, dtype=np.uint32)
        self.shape = words.shape
        if device:
            import torch
            self.t = torch.from_numpy(words.reshape(-1).view(np.int32).copy()).cuda()
            self.ptr = self.t.data_ptr()
        else:
            raw = np.zeros(words.size + 16, dtype=np.uint32)
            off = (-raw.ctypes.data % 64) // 4
            self.a = raw[off:off + words.size]
            self.a[:] = words.reshape(-1)
            self.ptr = self.a.ctypes.data

    def get(self):
        if self.device:
            return self.t.cpu().numpy().view(np.uint32).reshape(self.shape)
        return self.a.reshape(self.shape).copy()


def run_mode(product, sc, rays, mode, occluded=False, device=False, coherent=False):
    """Trace `rays` (RAYHIT_DTYPE for closest hit, RAY_DTYPE for occlusion) through entry point `mode`; returns the result array."""
    L = product.lib
    _sigs(L)
    ctx = product.context(coherent)
    r = rays.copy()
    n = len(r)
    names = RAY_NAMES if occluded else rt.RAYHIT_DTYPE.names
    nf = len(names)
    rec = 48 if occluded else 80
    if mode == "1":
        fn = L.rtcOccluded1 if occluded else L.rtcIntersect1
        for i in range(n):
            fn(sc, C.byref(ctx), r[i:i + 1].ctypes.data)
        return r
    if mode == "1M":
        fn = L.rtcOccluded1M if occluded else L.rtcIntersect1M
        if device:
            m = _Mem(r.view(np.uint32).reshape(n, rec // 4), True)
            fn(sc, C.byref(ctx), m.ptr, n, rec)
            return m.get().reshape(-1).view(r.dtype).copy()
        fn(sc, C.byref(ctx), r.ctypes.data, n, rec)
        return r
    if mode == "1Mp":
        fn = L.rtcOccluded1Mp if occluded else L.rtcIntersect1Mp
        if device:
            import torch
            m = _Mem(r.view(np.uint32).reshape(n, rec // 4), True)
            ptrs = torch.tensor([m.ptr + i * rec for i in range(n)], dtype=torch.int64).cuda()
            fn(sc, C.byref(ctx), ptrs.data_ptr(), n)
            return m.get().reshape(-1).view(r.dtype).copy()
        ptrs = (C.c_void_p * n)(*[r[i:i + 1].ctypes.data for i in range(n)])
        fn(sc, C.byref(ctx), ptrs, n)
        return r
    if mode in ("4", "8", "16"):
        w = int(mode)
        fn = getattr(L, ("rtcOccluded" if occluded else "rtcIntersect") + mode)
        for s in range(0, n, w):
            m = min(w, n - s)
            blk = _soa_block(r[s:s + m], names, w)
            vraw = np.zeros(w + 16, dtype=np.int32)                         # the reference loads the mask with aligned SIMD loads
            voff = (-vraw.ctypes.data % 64) // 4
            valid = vraw[voff:voff + w]
            valid[:m] = -1
            mem = _Mem(blk, device)
            fn(valid.ctypes.data, sc, C.byref(ctx), mem.ptr)
            part = r[s:s + m]
            _unsoa(part, mem.get(), names)
            r[s:s + m] = part
        return r
    if mode == "NM":
        N = 8
        fn = L.rtcOccludedNM if occluded else L.rtcIntersectNM
        blocks = (n + N - 1) // N
        full = _soa_block(r, names, blocks * N)                              # (nf, blocks*N)
        soa = np.ascontiguousarray(full.reshape(nf, blocks, N).transpose(1, 0, 2))   # (blocks, nf, N): packet-major
        mem = _Mem(soa, device)
        fn(sc, C.byref(ctx), mem.ptr, N, blocks, nf * N * 4)
        back = mem.get().reshape(blocks, nf, N).transpose(1, 0, 2).reshape(nf, blocks * N)
        _unsoa(r, back, names)
        return r
    if mode == "Np":
        fn = L.rtcOccludedNp if occluded else L.rtcIntersectNp
        mem = _Mem(_soa_block(r, names, n), device)
        ptrs = (C.c_void_p * nf)(*[mem.ptr + k * n * 4 for k in range(nf)])    # RTCRayNp / RTCRayHitNp: one pointer per field
        fn(sc, C.byref(ctx), ptrs, n)
        _unsoa(r, mem.get(), names)
        return r
    raise ValueError(mode)


def intersect_then_occluded_agree(product, sc, rays, mode, device=False):
    """The reference's VARIANT_INTERSECT_OCCLUDED check (rtcore_helpers.h:881-904): closest-hit and occlusion queries of one
    ray through the same entry point must agree on hit / no hit."""
    a = run_mode(product, sc, rays, mode, False, device)
    b = run_mode(product, sc, cases.fx.to_ray(rays), mode, True, device)
    hit = a["geomID"] != 0xFFFFFFFF
    occ = np.isneginf(b["tfar"])
    return a, b, np.array_equal(hit, occ)
