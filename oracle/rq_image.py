"""Independent reader of the product's flat BVH image (embree-aarch64_b200/csrc/rq_types.h) and evaluator of the reference's
SAH statistic on that tree.  TEST INFRASTRUCTURE ONLY (imported by tests/ and the bench's checker legs, never by the product).

The formula restated here is BVHNStatistics (reference kernels/bvh/bvh_statistics.cpp:41-160, bvh_statistics.h:36-38,99-101):
    sah = [ sum over inner nodes halfArea(node box) + sum over leaves halfArea(leaf box) * numBlocks ] / halfArea(root box)
where a node's own box is the box its PARENT stores for it (the root uses the scene bounds) and halfArea(d) = dx*(dy+dz)+dy*dz
(common/math/bbox.h halfArea).  In the product's tree a leaf is one slot of <= 3 triangles = one block.
"""
import numpy as np

MAGIC = 0x3276303032425152
NODE_DT = np.dtype([("p", "<f4", 3), ("e", "u1", 3), ("pad0", "u1"), ("childBase", "<u4"), ("triBase", "<u4"), ("masks", "<u4"),
                    ("pad1", "<u4"), ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8)), ("lo", "<f4", 3), ("hi", "<f4", 3),
                    ("parent", "<u4"), ("numTris", "<u4"), ("level", "<u4"), ("pad", "<u4", 3)])
assert NODE_DT.itemsize == 128
HEADER_DT = np.dtype([("magic", "<u8"), ("numNodes", "<u4"), ("numTris", "<u4"), ("depth", "<u4"), ("flags", "<u4"), ("lo", "<f4", 4),
                      ("hi", "<f4", 4), ("nodesOffset", "<u8"), ("trisOffset", "<u8"), ("totalBytes", "<u8"), ("sah", "<f8"),
                      ("metaOffset", "<u8"), ("vertsOffset", "<u8"), ("numVerts", "<u4"), ("layout", "<u4"), ("pad", "<u8", 2)])
assert HEADER_DT.itemsize == 128


def _half_area(dx, dy, dz):
    """float32, one rounding per operation (the product is compiled with -fmad=false)."""
    dx, dy, dz = (np.asarray(a, dtype=np.float32) for a in (dx, dy, dz))
    return (dx * (dy + dz).astype(np.float32)).astype(np.float32) + (dy * dz).astype(np.float32)


class Image:
    def __init__(self, raw):
        raw = np.frombuffer(bytes(raw), dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw.view(np.uint8).reshape(-1)
        self.header = raw[:128].view(HEADER_DT)[0]
        H = self.header
        assert int(H["magic"]) == MAGIC and int(H["totalBytes"]) == raw.size
        n, t = int(H["numNodes"]), int(H["numTris"])
        self.nodes = raw[int(H["nodesOffset"]):int(H["nodesOffset"]) + 128 * n].view(NODE_DT)
        self.compact = int(H["layout"]) == 1
        if self.compact:
            # RTC_SCENE_FLAG_COMPACT: 16-byte records (3 vertex-pool indices + primID), per-triangle geomID | quad flag, float4 pool
            rec = raw[int(H["trisOffset"]):int(H["trisOffset"]) + 16 * t].view("<u4").reshape(t, 4)
            meta = raw[int(H["metaOffset"]):int(H["metaOffset"]) + 4 * t].view("<u4")
            nv = int(H["numVerts"])
            pool = raw[int(H["vertsOffset"]):int(H["vertsOffset"]) + 16 * nv].view("<f4").reshape(nv, 4)
            assert (rec[:, :3] < max(nv, 1)).all(), "vertex index out of range"
            self.tri_words = rec
            self.verts = pool[rec[:, :3].astype(np.int64), :3]
            self.primID, self.geomID = rec[:, 3], meta & np.uint32(0x7FFFFFFF)
            self.pad = ((meta >> 31) << 30).astype(np.uint32)       # quad flag in the position check_structure expects
            return
        tr = raw[int(H["trisOffset"]):int(H["trisOffset"]) + 48 * t].view("<u4").reshape(t, 12).copy()
        odd = np.arange(t) & 1 == 1                      # odd records are stored rotated by 16 bytes (rq_types.h)
        tr[odd] = np.concatenate([tr[odd][:, 4:], tr[odd][:, :4]], axis=1)
        self.tri_words = tr
        self.verts = tr[:, :9].view("<f4").reshape(t, 3, 3)
        self.primID, self.geomID, self.pad = tr[:, 9], tr[:, 10], tr[:, 11]

    # ---- per-slot decoded data -------------------------------------------------------------------------------
    def slots(self):
        """Returns dict of arrays over (node, slot): inner / leaf masks, triangle counts, de-quantised extents."""
        N = self.nodes
        k = np.arange(8)
        imask = (N["masks"] >> 24)[:, None] >> k[None, :] & 1
        tbits = (N["masks"] & 0xFFFFFF)[:, None] >> (3 * k[None, :]) & 7
        ntri = (tbits & 1) + (tbits >> 1 & 1) + (tbits >> 2 & 1)
        step = (N["e"].astype(np.uint32) << 23).view("<f4")                 # (n, 3): 2^(e-127)
        dq = (N["qhi"].astype(np.float32) - N["qlo"].astype(np.float32)) * step[:, :, None]   # (n, 3, 8)
        lo = N["p"].astype(np.float64)[:, :, None] + N["qlo"].astype(np.float64) * step.astype(np.float64)[:, :, None]   # exact in fp64
        hi = N["p"].astype(np.float64)[:, :, None] + N["qhi"].astype(np.float64) * step.astype(np.float64)[:, :, None]
        return dict(inner=imask.astype(bool), leaf=ntri > 0, ntri=ntri, dq=dq.astype(np.float32), lo=lo, hi=hi)

    def sah(self):
        """(sah, inner term, leaf term, leaf-triangle term): the reference formula on the boxes traversal tests (de-quantised)."""
        S = self.slots()
        A = _half_area(S["dq"][:, 0, :], S["dq"][:, 1, :], S["dq"][:, 2, :]).astype(np.float64)
        H = self.header
        root = float(_half_area(H["hi"][0] - H["lo"][0], H["hi"][1] - H["lo"][1], H["hi"][2] - H["lo"][2]))
        n0 = self.nodes[0]
        rootBox = float(_half_area(*(n0["hi"] - n0["lo"]))) if (n0["masks"] != 0) else 0.0
        inner = float(A[S["inner"]].sum()) + rootBox                          # the root's own box counts as an inner node
        leaf = float(A[S["leaf"]].sum())                                      # one block per leaf slot (<= 3 triangles)
        leaf_tris = float((A * S["ntri"])[S["leaf"]].sum())
        if root <= 0:
            return 0.0, 0.0, 0.0, 0.0
        return (inner + leaf) / root, inner / root, leaf / root, leaf_tris / root

    def check_structure(self, expect_prims=None, presplit=False):
        """Structural invariants of a usable tree; raises AssertionError with the first violation.
        expect_prims: optional set-like array of (geomID << 33 | primID << 1 | flip bit) keys that must appear exactly once.
        presplit=True (RTC_BUILD_QUALITY_HIGH): a triangle may be stored once per clipped reference, so a leaf box need only
        overlap its triangle and the primitive set is compared without multiplicity."""
        N, S = self.nodes, self.slots()
        n, t = len(N), len(self.tri_words)
        ni = S["inner"].sum(1)
        nt = S["ntri"].sum(1)
        assert not (S["inner"] & S["leaf"]).any(), "slot both inner and leaf"
        has_i, has_t = ni > 0, nt > 0
        assert (N["childBase"][has_i] + ni[has_i] <= n).all(), "child range"
        assert (N["triBase"][has_t].astype(np.int64) + nt[has_t] <= t).all(), "triangle range"
        # every node except the root is referenced exactly once; every triangle exactly once
        ref = np.zeros(n, dtype=np.int64)
        idx = np.concatenate([np.arange(b, b + c) for b, c in zip(N["childBase"][has_i], ni[has_i])]) if has_i.any() else np.zeros(0, int)
        np.add.at(ref, idx, 1)
        assert ref[0] == 0 and (ref[1:] == 1).all(), "node reference counts"
        tref = np.zeros(t, dtype=np.int64)
        tidx = np.concatenate([np.arange(b, b + c) for b, c in zip(N["triBase"][has_t], nt[has_t])]) if has_t.any() else np.zeros(0, int)
        np.add.at(tref, tidx, 1)
        assert (tref == 1).all(), "triangle reference counts"
        # levels and depth
        if has_i.any():
            par = np.repeat(np.arange(n)[has_i], ni[has_i])
            assert (N["level"][idx] == N["level"][par] + 1).all(), "child level"
            assert (N["parent"][idx] == par).all(), "parent link"
        assert int(N["level"].max()) + 1 <= int(self.header["depth"]), "depth bound"
        # conservative quantisation: the decoded child box contains the child's exact box (inner) / its triangles (leaf)
        for k in range(8):
            m = S["inner"][:, k]
            if m.any():
                rank = np.array([bin(int(x) & ((1 << k) - 1)).count("1") for x in (N["masks"][m] >> 24)])
                ch = N["childBase"][m] + rank
                assert (S["lo"][m, :, k] <= N["lo"][ch]).all() and (S["hi"][m, :, k] >= N["hi"][ch]).all(), "inner child box not contained"
        tl = self.verts.min(1)
        th = self.verts.max(1)
        for k in range(8):
            m = S["leaf"][:, k]
            if not m.any():
                continue
            tv = (N["masks"][m] & 0xFFFFFF)
            first = N["triBase"][m] + np.array([bin(int(x) & ((1 << (3 * k)) - 1)).count("1") for x in tv])
            for j in range(3):
                mj = S["ntri"][m, k] > j
                ti = first[mj] + j
                sel = np.where(m)[0][mj]
                if presplit:
                    assert (S["lo"][sel, :, k] <= th[ti]).all() and (S["hi"][sel, :, k] >= tl[ti]).all(), "leaf box does not touch its triangle"
                else:
                    assert (S["lo"][sel, :, k] <= tl[ti]).all() and (S["hi"][sel, :, k] >= th[ti]).all(), "leaf triangle not contained"
        if expect_prims is not None:
            keys = (self.geomID.astype(np.uint64) << np.uint64(33)) | (self.primID.astype(np.uint64) << np.uint64(1)) | \
                   ((self.pad >> 30) & 1).astype(np.uint64)
            if presplit:
                keys = np.unique(keys)
            assert np.array_equal(np.sort(keys), np.sort(np.asarray(expect_prims, dtype=np.uint64))), "primitive set"
        return True


def fetch(product, scene):
    """Host copy of the committed image of `scene` (rtcxGetSceneImage + rtcxCopySceneImage)."""
    import ctypes as C
    nbytes = C.c_size_t(0)
    product.lib.rtcxGetSceneImage(scene, C.byref(nbytes))
    buf = np.zeros(nbytes.value, dtype=np.uint8)
    product.lib.rtcxCopySceneImage(scene, buf.ctypes.data, nbytes.value)
    return Image(buf)
