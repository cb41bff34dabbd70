"""TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/_ref/liblexp_ref.so: the reference's own classes
(CostVolumeEnergy, NaiveStereoEnergy, FastGuidedImageFilter<double>, LayerManager, RandomProposer) compiled from
/root/reference by oracle/build_ref.py.  Used to pin the restated oracles and to generate tests/golden/ref_*.npz."""
import ctypes as C
import os

import numpy as np

from . import build_ref

_lib = None


def available():
    return build_ref.available()


def lib():
    global _lib
    if _lib is None:
        path = build_ref.build()
        if path is None or not os.path.exists(path):
            raise RuntimeError("oracle/_ref/liblexp_ref.so is absent and the reference sources are not here to build it")
        L = C.CDLL(path)
        vp, ip, fp, dp, u8p = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_ubyte)
        L.ref_last_error.restype = C.c_char_p
        L.ref_max_threads.restype = C.c_int
        L.ref_create.restype = vp
        L.ref_create.argtypes = [C.c_int] * 4 + [u8p, u8p, fp, fp, C.c_int] + [C.c_float] * 6
        L.ref_destroy.argtypes = [vp]
        L.ref_unary.argtypes = [vp, ip, ip, fp, C.c_int, C.c_int, fp]
        L.ref_unary_group.argtypes = [vp, C.c_int, ip, ip, fp, C.c_int, C.c_int, C.c_int, fp, C.c_int]
        L.ref_stats.argtypes = [vp, C.c_int, dp]
        L.ref_exi.argtypes = [vp, C.c_int, fp]
        L.ref_valid_mask.argtypes = [vp, fp, ip, u8p]
        L.ref_rng_seed.argtypes = [C.c_uint64]
        L.ref_rng_state.restype = C.c_uint64
        L.ref_create_random_label.argtypes = [vp, C.c_int, C.c_int, fp]
        L.ref_random_proposals.argtypes = [fp, C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_float, C.c_float, fp]
        L.ref_plane_normal.argtypes = [fp, fp]
        L.ref_create_plane.argtypes = [fp] + [C.c_float] * 4 + [fp]
        L.ref_layer_create.restype = vp
        L.ref_layer_create.argtypes = [C.c_int] * 4
        L.ref_layer_destroy.argtypes = [vp]
        L.ref_layer_counts.argtypes = [vp, ip]
        L.ref_layer_rects.argtypes = [vp, ip, ip, ip]
        L.ref_layer_group.argtypes = [vp, C.c_int, ip]
        L.ref_layer_group.restype = C.c_int
        u64p = C.POINTER(C.c_uint64)
        L.ref_pm_group.argtypes = [vp, C.c_int, C.c_int, ip, ip, ip, C.c_int, ip, ip, C.c_int, C.c_float, C.c_float, fp, C.c_int, u64p, C.c_int,
                                   fp, fp, fp, ip, C.c_int]
        L.ref_pm_init.argtypes = [vp, C.c_int, C.c_int, ip, fp, C.c_int, fp, fp, C.c_int]
        L.ref_set_smoothness.argtypes = [vp] + [C.c_float] * 4
        L.ref_smooth_coeff.argtypes = [vp, C.c_int, fp]
        L.ref_smooth_terms_expansion.argtypes = [vp, C.c_int, fp, fp, ip, fp]
        L.ref_gc_group.argtypes = [vp, C.c_int, C.c_int, ip, ip, ip, C.c_int, ip, ip, C.c_int, fp, C.c_int, u64p, C.c_int, fp, fp, fp, ip, dp, C.c_int]
        L.ref_smoothness_cost.argtypes = [vp, C.c_int, fp]
        L.ref_smoothness_cost.restype = C.c_double
        L.ref_save_pfm.argtypes = [C.c_char_p, fp, C.c_int, C.c_int]
        L.ref_read_pfm.argtypes = [C.c_char_p, fp, C.c_int, C.c_int]
        L.ref_load_acrt.argtypes = [C.c_char_p, fp, C.c_int, C.c_int, C.c_int]
        L.ref_disparities.argtypes = [vp, fp, fp]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i4(r):
    return np.ascontiguousarray(np.asarray(r, dtype=np.int32).reshape(-1))


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class RefEnergy:
    """kind 0 = CostVolumeEnergy(imL, imR, volL, volR, Parameters(windR, "GF", eps), MAX, MIN); kind 1 = NaiveStereoEnergy."""

    def __init__(self, imL, imR, volL=None, volR=None, windR=20, eps=1e-4, th_col=0.5, th_grad=2.0, alpha=0.9, max_disp=None, min_disp=0.0, kind=0):
        L = lib()
        self.imL = np.ascontiguousarray(imL, dtype=np.uint8)
        self.imR = np.ascontiguousarray(imR, dtype=np.uint8)
        self.H, self.W = self.imL.shape[:2]
        self.kind = kind
        if kind == 0:
            self.volL, self.volR = _f(volL), _f(volR)  # borrowed by the C++ side: kept alive here
            self.D = self.volL.shape[0]
            vl, vr = _p(self.volL, C.c_float), _p(self.volR, C.c_float)
        else:
            self.D, vl, vr = 0, None, None
        self.max_disp, self.min_disp = float(max_disp), float(min_disp)
        self.h = L.ref_create(kind, self.H, self.W, self.D, _p(self.imL, C.c_ubyte), _p(self.imR, C.c_ubyte), vl, vr, int(windR),
                              float(eps), float(th_col), float(th_grad), float(alpha), float(max_disp), float(min_disp))
        if not self.h:
            raise RuntimeError(L.ref_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None

    __del__ = close

    def unary(self, filter_rect, target_rect, plane, mode=0, with_check=True, fill=np.nan):
        """Returns the filterRect-sized cost view after the call (untouched pixels keep `fill`)."""
        fr, tr, pl = _i4(filter_rect), _i4(target_rect), _f(plane)
        out = np.full((fr[3], fr[2]), fill, np.float32)
        if lib().ref_unary(self.h, _p(fr, C.c_int), _p(tr, C.c_int), _p(pl, C.c_float), int(mode), int(with_check), _p(out, C.c_float)):
            raise RuntimeError(lib().ref_last_error().decode())
        return out

    def unary_target(self, filter_rect, target_rect, plane, mode=0, with_check=True):
        fx, fy = filter_rect[0], filter_rect[1]
        tx, ty, tw, th = target_rect
        return self.unary(filter_rect, target_rect, plane, mode, with_check)[ty - fy:ty - fy + th, tx - fx:tx - fx + tw].copy()

    def unary_group(self, filter_rects, target_rects, planes, mode=0, with_check=True, cost_image=None, nthreads=0):
        """planes: [n][K][4].  Returns the H x W cost image after the last proposal of every cell."""
        fr, tr = _i4(filter_rects), _i4(target_rects)
        pl = _f(planes)
        n, K = pl.shape[0], pl.shape[1]
        if cost_image is None:
            cost_image = np.zeros((self.H, self.W), np.float32)
        if lib().ref_unary_group(self.h, n, _p(fr, C.c_int), _p(tr, C.c_int), _p(pl, C.c_float), K, int(mode), int(with_check), _p(cost_image, C.c_float), int(nthreads)):
            raise RuntimeError(lib().ref_last_error().decode())
        return cost_image

    def stats(self, mode=0):
        out = np.empty((9, self.H, self.W), np.float64)
        if lib().ref_stats(self.h, int(mode), _p(out, C.c_double)):
            raise RuntimeError(lib().ref_last_error().decode())
        return out

    def exi(self, mode=0):
        out = np.empty((self.H, self.W, 4), np.float32)
        if lib().ref_exi(self.h, int(mode), _p(out, C.c_float)):
            raise RuntimeError(lib().ref_last_error().decode())
        return out

    def valid_mask(self, plane, rect):
        r, pl = _i4(rect), _f(plane)
        out = np.empty((r[3], r[2]), np.uint8)
        if lib().ref_valid_mask(self.h, _p(pl, C.c_float), _p(r, C.c_int), _p(out, C.c_ubyte)):
            raise RuntimeError(lib().ref_last_error().decode())
        return out

    def pm_group(self, units, shareds, filts, proposers, outer_iter, states, cur_cost, cur_label, list_planes=None, mode=0, nthreads=0):
        """FastGCStereo.h:30-61 with doGC == false for the cells of one disjoint group, with the reference's own proposers.
        proposers: [(kind, K)] with kind 0 = replayed plane list, 1 = ExpansionProposer, 2 = RandomProposer; states: uint64 [n][max_steps]
        cv::RNG states set before each proposal; cur_cost [H][W] / cur_label [H][W][4] float32, updated in place.
        Returns (planes [n][max_steps][4], steps [n])."""
        n = len(units)
        u, sh, fl = _i4(units), _i4(shareds), _i4(filts)
        kinds = np.ascontiguousarray([k for k, _ in proposers], dtype=np.int32)
        Ks = np.ascontiguousarray([K for _, K in proposers], dtype=np.int32)
        st = np.ascontiguousarray(states, dtype=np.uint64)
        assert st.ndim == 2 and st.shape[0] == n
        max_steps = st.shape[1]
        list_steps = int(sum(K for k, K in proposers if k == 0))
        lp = _f(list_planes if list_planes is not None else np.zeros((n, max(list_steps, 1), 4), np.float32))
        assert list_steps == 0 or lp.shape == (n, list_steps, 4)
        assert cur_cost.dtype == np.float32 and cur_cost.flags.c_contiguous and cur_label.dtype == np.float32 and cur_label.flags.c_contiguous
        out = np.zeros((n, max_steps, 4), np.float32)
        steps = np.zeros(n, np.int32)
        rc = lib().ref_pm_group(self.h, mode, n, _p(u, C.c_int), _p(sh, C.c_int), _p(fl, C.c_int), len(proposers), _p(kinds, C.c_int),
                                _p(Ks, C.c_int), int(outer_iter), float(self.max_disp), float(self.min_disp), _p(lp, C.c_float), list_steps,
                                _p(st, C.c_uint64), max_steps, _p(cur_cost, C.c_float), _p(cur_label, C.c_float), _p(out, C.c_float),
                                _p(steps, C.c_int), nthreads)
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())
        return out, steps

    def pm_init(self, units, labels, windR, cur_cost, cur_label, mode=0, nthreads=0):
        """initCurrentFast (FastGCStereo.h:101-113) with the given label per unit region."""
        u, lb = _i4(units), _f(labels)
        rc = lib().ref_pm_init(self.h, mode, len(units), _p(u, C.c_int), _p(lb, C.c_float), int(windR), _p(cur_cost, C.c_float),
                               _p(cur_label, C.c_float), nthreads)
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())

    # ---- pairwise terms / graph-cut move ----
    def set_smoothness(self, lam=1.0, omega=10.0, th_smooth=1.0, epsilon=0.01):
        """params.lambda / omega / th_smooth / epsilon, then the reference's initSmoothnessCoeff() (StereoEnergy.h:131-163)."""
        if lib().ref_set_smoothness(self.h, float(lam), float(omega), float(th_smooth), float(epsilon)):
            raise RuntimeError(lib().ref_last_error().decode())

    def smooth_coeff(self, mode=0):
        """smoothnessCoeff[mode][k] without the margin: float32 [8][H][W] (k in the order of StereoEnergy::NB_*)."""
        out = np.zeros((8, self.H, self.W), np.float32)
        if lib().ref_smooth_coeff(self.h, int(mode), _p(out, C.c_float)) < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return out

    def smooth_terms_expansion(self, labeling, plane, region, mode=0):
        """computeSmoothnessTermsExpansion(.., onlyForward = true) as expansionMoveBK calls it (FastGCStereo.h:422):
        returns (cost00, cost01, cost10), each float32 [8][rh][rw] (only the forward neighbours 1, 3, 6, 7 are filled)."""
        lab, pl, rg = _f(labeling), _f(plane), _i4(region)
        assert lab.shape == (self.H, self.W, 4)
        out = np.zeros((3, 8, rg[3], rg[2]), np.float32)
        if lib().ref_smooth_terms_expansion(self.h, int(mode), _p(lab, C.c_float), _p(pl, C.c_float), _p(rg, C.c_int), _p(out, C.c_float)):
            raise RuntimeError(lib().ref_last_error().decode())
        return out[0], out[1], out[2]

    def smoothness_cost(self, labeling, mode=0):
        lab = _f(labeling)
        return float(lib().ref_smoothness_cost(self.h, int(mode), _p(lab, C.c_float)))

    def gc_group(self, units, shareds, filts, proposers, outer_iter, states, cur_cost, cur_label, list_planes=None, mode=0, nthreads=0):
        """FastGCStereo.h:30-61 with doGC == true for the cells of one disjoint group: the reference's own proposers, energy and
        FastGCStereo::expansionMoveBK (over oracle/maxflow/graph.h).  Arguments as pm_group.  Returns (planes, steps, flows [n][max_steps])."""
        n = len(units)
        u, sh, fl = _i4(units), _i4(shareds), _i4(filts)
        kinds = np.ascontiguousarray([k for k, _ in proposers], dtype=np.int32)
        Ks = np.ascontiguousarray([K for _, K in proposers], dtype=np.int32)
        st = np.ascontiguousarray(states, dtype=np.uint64)
        assert st.ndim == 2 and st.shape[0] == n
        max_steps = st.shape[1]
        list_steps = int(sum(K for k, K in proposers if k == 0))
        lp = _f(list_planes if list_planes is not None else np.zeros((n, max(list_steps, 1), 4), np.float32))
        assert list_steps == 0 or lp.shape == (n, list_steps, 4)
        assert cur_cost.dtype == np.float32 and cur_cost.flags.c_contiguous and cur_label.dtype == np.float32 and cur_label.flags.c_contiguous
        out = np.zeros((n, max_steps, 4), np.float32)
        steps = np.zeros(n, np.int32)
        flows = np.zeros((n, max_steps), np.float64)
        rc = lib().ref_gc_group(self.h, mode, n, _p(u, C.c_int), _p(sh, C.c_int), _p(fl, C.c_int), len(proposers), _p(kinds, C.c_int),
                                _p(Ks, C.c_int), int(outer_iter), _p(lp, C.c_float), list_steps, _p(st, C.c_uint64), max_steps,
                                _p(cur_cost, C.c_float), _p(cur_label, C.c_float), _p(out, C.c_float), _p(steps, C.c_int), _p(flows, C.c_double), nthreads)
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())
        return out, steps, flows

    def disparities(self, labeling):
        """StereoEnergy::computeDisparities (StereoEnergy.h:269-272)."""
        lab = _f(labeling)
        out = np.empty((self.H, self.W), np.float32)
        if lib().ref_disparities(self.h, _p(lab, C.c_float), _p(out, C.c_float)):
            raise RuntimeError(lib().ref_last_error().decode())
        return out

    def create_random_label(self, x, y):
        out = np.empty(4, np.float32)
        lib().ref_create_random_label(self.h, int(x), int(y), _p(out, C.c_float))
        return out


def rng_seed(s):
    lib().ref_rng_seed(int(s))


def rng_state():
    return int(lib().ref_rng_state())


def random_proposals(labeling, unit, outer_iter, K, max_disp, min_disp=0.0):
    lab = _f(labeling)
    H, W = lab.shape[:2]
    out = np.empty((K, 4), np.float32)
    u = _i4(unit)
    n = lib().ref_random_proposals(_p(lab, C.c_float), H, W, _p(u, C.c_int), int(outer_iter), int(K), float(max_disp), float(min_disp), _p(out, C.c_float))
    return out[:n]


def plane_normal(plane):
    out = np.empty(3, np.float32)
    pl = _f(plane)
    lib().ref_plane_normal(_p(pl, C.c_float), _p(out, C.c_float))
    return out


def create_plane(n, z, x, y, v=0.0):
    out = np.empty(4, np.float32)
    nn = _f(n)
    lib().ref_create_plane(_p(nn, C.c_float), float(z), float(x), float(y), float(v), _p(out, C.c_float))
    return out


def layer(W, H, windR, unit):
    """LayerManager(W, H, windR, 0).addLayer(unit) -> the same dict shape as lexp_oracle.make_layer."""
    L = lib()
    h = L.ref_layer_create(int(W), int(H), int(windR), int(unit))
    try:
        cnt = np.zeros(4, np.int32)
        L.ref_layer_counts(h, _p(cnt, C.c_int))
        hb, wb, n, ng = (int(v) for v in cnt)
        u, s, f = (np.zeros((n, 4), np.int32) for _ in range(3))
        L.ref_layer_rects(h, _p(u, C.c_int), _p(s, C.c_int), _p(f, C.c_int))
        groups = []
        for g in range(ng):
            m = L.ref_layer_group(h, g, None)
            idx = np.zeros(m, np.int32)
            L.ref_layer_group(h, g, _p(idx, C.c_int))
            groups.append([int(i) for i in idx])
        as_t = lambda a: [tuple(int(v) for v in r) for r in a]
        return dict(heightBlocks=hb, widthBlocks=wb, unitSize=int(unit), unit=as_t(u), shared=as_t(s), filter=as_t(f), groups=groups)
    finally:
        L.ref_layer_destroy(h)


# ---- the cv:: layer's own primitives (held against cv2 in tests/test_ref_pin.py) --------------------------------------
def shim_box_sum(X, R):
    X = np.ascontiguousarray(X)
    assert X.dtype in (np.float32, np.float64) and X.ndim == 2
    out = np.empty_like(X)
    L = lib()
    L.shim_box_sum.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    if L.shim_box_sum(X.ctypes.data, int(X.dtype == np.float64), X.shape[0], X.shape[1], int(R), out.ctypes.data):
        raise RuntimeError(L.ref_last_error().decode())
    return out


def shim_get_affine(src_pts, dst_pts):
    s, d = _f(src_pts).reshape(-1), _f(dst_pts).reshape(-1)
    M = np.empty(6, np.float64)
    L = lib()
    L.shim_get_affine.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_double)]
    L.shim_get_affine(_p(s, C.c_float), _p(d, C.c_float), _p(M, C.c_double))
    return M.reshape(2, 3)


def shim_warp_affine(src, M, w, h):
    src = _f(src)
    H, W = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    M = np.ascontiguousarray(np.asarray(M, np.float64).reshape(-1))
    out = np.empty((h, w) + src.shape[2:], np.float32)
    L = lib()
    L.shim_warp_affine.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_float)]
    if L.shim_warp_affine(_p(src, C.c_float), H, W, cn, _p(M, C.c_double), int(h), int(w), _p(out, C.c_float)):
        raise RuntimeError(L.ref_last_error().decode())
    return out


def shim_bgr2gray(src):
    src = _f(src)
    out = np.empty(src.shape[:2], np.float32)
    L = lib()
    L.shim_bgr2gray.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float)]
    if L.shim_bgr2gray(_p(src, C.c_float), src.shape[0], src.shape[1], _p(out, C.c_float)):
        raise RuntimeError(L.ref_last_error().decode())
    return out


def shim_sobel_x(src, scale):
    src = _f(src)
    out = np.empty(src.shape, np.float32)
    L = lib()
    L.shim_sobel_x.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_double, C.POINTER(C.c_float)]
    if L.shim_sobel_x(_p(src, C.c_float), src.shape[0], src.shape[1], float(scale), _p(out, C.c_float)):
        raise RuntimeError(L.ref_last_error().decode())
    return out


def shim_grid_mincut(tr, cap):
    """The BK stand-in oracle/maxflow/graph.h (what the compiled reference's expansionMoveBK runs on) on a grid given as arrays."""
    tr, cap = _f(tr), _f(cap)
    h, w = tr.shape
    mask = np.empty((h, w), np.uint8)
    L = lib()
    L.shim_grid_mincut.restype = C.c_double
    L.shim_grid_mincut.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_ubyte)]
    f = L.shim_grid_mincut(w, h, _p(tr, C.c_float), _p(cap, C.c_float), _p(mask, C.c_ubyte))
    return mask.astype(bool), float(f)


def save_pfm(path, img):
    """cvutils::io::save_pfm_file of the reference (Utilities.hpp:84-137)."""
    a = _f(img)
    if lib().ref_save_pfm(str(path).encode(), _p(a, C.c_float), a.shape[0], a.shape[1]):
        raise RuntimeError(lib().ref_last_error().decode())


def read_pfm(path, H, W):
    out = np.empty((H, W), np.float32)
    if lib().ref_read_pfm(str(path).encode(), _p(out, C.c_float), H, W):
        raise RuntimeError(lib().ref_last_error().decode())
    return out


def load_acrt(path, D, H, W):
    """cvutils::io::loadMatBinary(path, vol, false) as main.cpp:353-358 reads im0.acrt."""
    out = np.empty((D, H, W), np.float32)
    if lib().ref_load_acrt(str(path).encode(), _p(out, C.c_float), D, H, W):
        raise RuntimeError("loadMatBinary failed")
    return out
