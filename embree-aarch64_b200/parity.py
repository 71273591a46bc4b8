"""Comparison of two traced ray streams under the north-star tolerances (BASELINE.json):
hit/miss agreement >= 99.99 %, geomID/primID exact on agreed hits, t within 1e-5 relative,
u/v within 1e-4 absolute; disagreements are only tolerated for rays within eps of a shared
edge or vertex (or on exactly coincident geometry, where the reference's own answer depends on
its BVH's test order: triangle_intersector_moeller.h:95, vfloat8_avx.h:726-731).
Used by tests and bench only (a checker, never on the product path).
"""
import numpy as np

INVALID = 0xFFFFFFFF


def compare_closest(ours, ref, edge_eps=1e-4, t_rel=1e-5, uv_abs=1e-4):
    """ours/ref: RTCRayHit streams traced from identical inputs.  Returns a dict of counts."""
    n = len(ours)
    ho, hr = ours["geomID"] != INVALID, ref["geomID"] != INVALID
    both = ho & hr
    same_id = both & (ours["geomID"] == ref["geomID"]) & (ours["primID"] == ref["primID"]) & (ours["instID"] == ref["instID"])
    # barycentric distance to the nearest edge, from whichever side reported a hit
    def edge_dist(r):
        return np.minimum(np.minimum(r["u"], r["v"]), 1.0 - r["u"] - r["v"])
    near_edge = (np.where(ho, np.abs(edge_dist(ours)), 1.0) < edge_eps) | (np.where(hr, np.abs(edge_dist(ref)), 1.0) < edge_eps)
    hitmiss_dis = ho != hr
    id_dis = both & ~same_id
    with np.errstate(invalid="ignore", divide="ignore"):
        t_err = np.abs(ours["tfar"] - ref["tfar"]) / np.maximum(np.abs(ref["tfar"]), 1e-30)
    same_t = t_err <= 10 * t_rel                       # "same surface point" for id disagreements (coplanar / shared edge)
    du, dv = np.abs(ours["u"] - ref["u"]), np.abs(ours["v"] - ref["v"])
    ngo = np.stack([ours["Ng_x"], ours["Ng_y"], ours["Ng_z"]], 1)
    ngr = np.stack([ref["Ng_x"], ref["Ng_y"], ref["Ng_z"]], 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        ng_err = np.linalg.norm(ngo - ngr, axis=1) / np.maximum(np.linalg.norm(ngr, axis=1), 1e-30)
    res = {
        "rays": int(n),
        "hits_ours": int(ho.sum()), "hits_ref": int(hr.sum()),
        "hitmiss_disagree": int(hitmiss_dis.sum()),
        "hitmiss_disagree_not_near_edge": int((hitmiss_dis & ~near_edge).sum()),
        "id_disagree": int(id_dis.sum()),
        "id_disagree_unexplained": int((id_dis & ~(near_edge | same_t)).sum()),
        "agreement": float(1.0 - hitmiss_dis.sum() / max(n, 1)),
        "max_t_rel": float(t_err[same_id].max()) if same_id.any() else 0.0,
        "max_uv_abs": float(max(du[same_id].max(), dv[same_id].max())) if same_id.any() else 0.0,
        "max_ng_rel": float(ng_err[same_id].max()) if same_id.any() else 0.0,
        "t_out_of_tol": int((t_err[same_id] > t_rel).sum()),
        "uv_out_of_tol": int(((du[same_id] > uv_abs) | (dv[same_id] > uv_abs)).sum()),
        "untouched_miss_ok": bool(np.array_equal(ours["tfar"][~ho & ~hr].view(np.uint32), ref["tfar"][~ho & ~hr].view(np.uint32))),
    }
    res["pass"] = bool(res["agreement"] >= 0.9999 and res["hitmiss_disagree_not_near_edge"] == 0 and
                       res["id_disagree_unexplained"] == 0 and res["t_out_of_tol"] == 0 and
                       res["uv_out_of_tol"] == 0 and res["untouched_miss_ok"])
    return res


def compare_occluded(ours, ref, closest_ref=None, edge_eps=1e-4):
    """ours/ref: RTCRay streams after rtcOccluded1M.  Occluded <=> tfar == -inf; everything else untouched."""
    oo, orf = np.isneginf(ours["tfar"]), np.isneginf(ref["tfar"])
    dis = oo != orf
    res = {"rays": int(len(ours)), "occluded_ours": int(oo.sum()), "occluded_ref": int(orf.sum()),
           "disagree": int(dis.sum()), "agreement": float(1.0 - dis.sum() / max(len(ours), 1))}
    same = ~oo & ~orf
    res["untouched_ok"] = bool(np.array_equal(ours["tfar"][same].view(np.uint32), ref["tfar"][same].view(np.uint32)))
    res["pass"] = bool(res["agreement"] >= 0.9999 and res["untouched_ok"])
    return res
