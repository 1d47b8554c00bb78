"""CPU oracle (TEST INFRASTRUCTURE ONLY) for row N4 of SURVEY.md section 8(f): cv::StereoBM::compute as the reference's
MatcherOpenCVBlock calls it (/root/reference/src/stereoMatcher/matcherOpenCVBlock.cpp:13-20: StereoBM::create(64, 9),
compute(left, right, CV_16S disparity); setters :52-110).  The arithmetic lives in OpenCV calib3d (absent from /root/reference;
oracle version = cv2 4.13.0): a numpy restatement of the published algorithm (PREFILTER_XSOBEL, SAD block matching with
replicate-clamped windows, texture threshold, uniqueness, x16 sub-pixel, valid-ROI mask, optional speckle filter), pinned
against cv2 by tests/test_bm_oracle.py (live) and tests/golden/bm_small.npz.  disp12MaxDiff >= 0 is not restated (the reference
never sets it for this matcher: generate_disparity.cpp:241-261 does not forward it).  Only tests/ may import this module."""
from __future__ import annotations

import numpy as np


def prefilter_xsobel(src: np.ndarray, cap: int) -> np.ndarray:
    """prefilterXSobel: clip(sobel_x, -cap, cap) + cap with reflect-101 rows; columns 0 and W-1 = cap; rows are produced in
    pairs, so the last row of an odd-height image (and the only row of a 1-row image) is all cap."""
    H, W = src.shape
    s = src.astype(np.int32)
    dst = np.full((H, W), cap, np.uint8)
    if W < 3:
        return dst
    rows = np.arange(H)
    up = np.where(rows > 0, rows - 1, np.minimum(rows + 1, H - 1))
    dn = np.where(rows < H - 1, rows + 1, np.maximum(rows - 1, 0))
    dx = s[:, 2:] - s[:, :-2]
    g = dx[up] + 2 * dx + dx[dn]
    v = (np.clip(g, -cap, cap) + cap).astype(np.uint8)
    npair = (H - 1 + 1) // 2 * 2 if H > 1 else 0          # rows covered by the pair loop: y < H-1, y += 2
    npair = min(npair, H)
    dst[:npair, 1:W - 1] = v[:npair]
    return dst


def compute(left: np.ndarray, right: np.ndarray, numDisparities: int = 64, blockSize: int = 9, minDisparity: int = 0,
            preFilterCap: int = 31, textureThreshold: int = 10, uniquenessRatio: int = 15, speckleWindowSize: int = 0,
            speckleRange: int = 0) -> np.ndarray:
    H, W = left.shape
    ndisp, wsz, mindisp, cap = numDisparities, blockSize, minDisparity, preFilterCap
    w2 = wsz // 2
    FILTERED = (mindisp - 1) << 4
    disp = np.full((H, W), FILTERED, np.int16)
    lofs = max(ndisp - 1 + mindisp, 0)
    rofs = -min(ndisp - 1 + mindisp, 0)
    width1 = W - rofs - ndisp + 1
    if lofs >= W or rofs >= W or width1 < 1:
        return disp
    L = prefilter_xsobel(left, cap).astype(np.int32)
    R = prefilter_xsobel(right, cap).astype(np.int32)
    xx = np.arange(-w2 - 1, width1 + w2)                      # window columns, one extra on the left for the prefix sum
    lc = np.clip(xx, -lofs, W - lofs - 1) + lofs
    rc = np.clip(xx, -rofs, W - rofs - ndisp) + rofs
    d = np.arange(ndisp)
    ad = np.abs(L[:, lc][:, :, None] - R[:, rc[:, None] + d[None, :]]).astype(np.int64)     # H x nx x ndisp
    tx = np.abs(L[:, lc] - cap).astype(np.int64)                                           # H x nx
    cs = np.cumsum(ad, axis=1)
    hs = cs[:, wsz:wsz + width1] - cs[:, 0:width1]                                          # window x-w2 .. x+w2
    ct = np.cumsum(tx, axis=1)
    ht = ct[:, wsz:wsz + width1] - ct[:, 0:width1]
    rows = np.clip(np.arange(-w2 - 1, H + w2), 0, H - 1)
    vs = np.cumsum(hs[rows], axis=0)
    sad = vs[wsz:wsz + H] - vs[0:H]                                                        # H x width1 x ndisp
    vt = np.cumsum(ht[rows], axis=0)
    tsum = vt[wsz:wsz + H] - vt[0:H]
    mind = np.argmin(sad, axis=2)                                                          # first minimum
    minsad = np.take_along_axis(sad, mind[:, :, None], 2)[:, :, 0]
    ok = tsum >= textureThreshold
    if uniquenessRatio > 0:
        thresh = minsad + (minsad * uniquenessRatio // 100)
        far = (d[None, None, :] < (mind - 1)[:, :, None]) | (d[None, None, :] > (mind + 1)[:, :, None])
        ok &= ~np.any(far & (sad <= thresh[:, :, None]), axis=2)
    im = np.where(mind == 0, 1, mind - 1)
    ip = np.where(mind == ndisp - 1, ndisp - 2, mind + 1) if ndisp > 1 else np.zeros_like(mind)
    n = np.take_along_axis(sad, im[:, :, None], 2)[:, :, 0]
    p = np.take_along_axis(sad, ip[:, :, None], 2)[:, :, 0]
    den = p + n - 2 * minsad + np.abs(p - n)
    num = (p - n) * 256
    q = np.where(den != 0, np.sign(num) * (np.abs(num) // np.maximum(den, 1)), 0)          # C division: toward zero
    val = ((ndisp - mind - 1 + mindisp) * 256 + q + 15) >> 4
    cols = lofs + np.arange(width1)
    keep = cols < W
    out = np.where(ok, val, FILTERED).astype(np.int16)
    disp[:, cols[keep]] = out[:, keep]
    # valid disparity ROI (getValidDisparityROI with full-image ROIs)
    maxD = mindisp + ndisp - 1
    xmin, xmax, ymin, ymax = max(0, maxD) + w2, min(W, W - min(mindisp, 0) * 0) - w2, w2, H - w2
    mask = np.zeros((H, W), bool)
    if xmax > xmin and ymax > ymin:
        mask[ymin:ymax, max(xmin, 0):xmax] = True
    disp[~mask] = FILTERED
    # cv::StereoBM walks width1 columns starting at lofs, i.e. minDisparity pixels past the end of every row (minDisparity > 0):
    # the spill of row y lands in the first pixels of row y + 1.  Inside the ROI rows the mask above wipes it; the spill of the
    # LAST computed row (H - w2 - 1) stays in row H - w2.  Reproduced because the reference's launch default has minDisparity > 0.
    ys = H - w2
    if mindisp > 0 and w2 >= 1 and 1 <= ys < H and xmax > xmin and ymax > ymin:
        k = min(width1 - (W - lofs), W)
        if k > 0:
            disp[ys, :k] = out[ys - 1, W - lofs:W - lofs + k]
    if speckleRange >= 0 and speckleWindowSize > 0:
        from oracle import oracle
        disp = oracle.filter_speckles(disp, FILTERED, speckleWindowSize, speckleRange)
    return disp


def overflow_mask(shape, blockSize: int, minDisparity: int) -> np.ndarray:
    """Pixels where cv::StereoBM leaves values it wrote past the end of the previous row (minDisparity > 0): the first
    minDisparity columns of the first row below the valid ROI.  compute() reproduces them since round 2; kept for tools that
    compare against older fixtures."""
    m = np.zeros(shape, bool)
    if minDisparity > 0:
        y = shape[0] - blockSize // 2
        if 0 <= y < shape[0]:
            m[y, :minDisparity] = True
    return m
