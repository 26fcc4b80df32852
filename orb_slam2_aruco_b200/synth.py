"""Deterministic synthetic gray frames for tests and bench.

numpy only, integer arithmetic wherever a rounding could differ between machines.  Frame `i` of a
run uses numpy.random.Generator(PCG64(20260000 + i)).

Scene: blurred-noise background stretched to [30,225] + 150 random filled rectangles / triangles
(sides 8-60 px at 640 wide, scaled with the width) so that every pyramid level can fill its ORB
quota; optionally ~20 square fiducial markers (rendered like reference
Thirdparty/aruco/aruco/dictionary.cpp:254-284: bit 0 at the bottom-right cell, one-cell black
border, plus a one-cell white quiet zone) with in-plane rotation and mild perspective on a jittered
5x4 grid (sides 50-76 px at 640 wide so that rotated quiet zones of neighbours rarely overlap); finally a sigma=0.7 blur and sigma=2 Gaussian noise.
"""
import os
import re

import numpy as np

_DICTS = None


def dictionaries():
    """{name: (nbits, tau, [codes])} parsed from csrc/aruco_dicts.inc (the single source of the tables)."""
    global _DICTS
    if _DICTS is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "aruco_dicts.inc")
        txt = open(path).read()
        d = {}
        for m in re.finditer(r"DICT_BEGIN\((\w+), (\d+), (\d+), (\d+)\)\n(.*?)DICT_END", txt, re.S):
            codes = [int(c, 16) for c in re.findall(r"C\((0x[0-9a-f]+)\)", m.group(5))]
            assert len(codes) == int(m.group(4))
            d[m.group(1)] = (int(m.group(2)), int(m.group(3)), codes)
        _DICTS = d
    return _DICTS


def marker_cells(dict_name, marker_id):
    """(n+2)x(n+2) 0/1 cell matrix of a marker incl. its black border (dictionary.cpp:254-284)."""
    nbits, _, codes = dictionaries()[dict_name]
    n = int(round(nbits ** 0.5))
    code = codes[marker_id]
    cells = np.zeros((n + 2, n + 2), np.uint8)
    b = 0
    for y in range(n - 1, -1, -1):
        for x in range(n - 1, -1, -1):
            cells[1 + y, 1 + x] = (code >> b) & 1
            b += 1
    return cells


def _gauss_kernel_q14(sigma):
    r = max(1, int(np.ceil(3 * sigma)))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-x * x / (2 * sigma * sigma))
    k = np.floor(k / k.sum() * 16384 + 0.5).astype(np.int64)
    k[r] += 16384 - k.sum()
    return k, r


def _blur_int(img, sigma):
    """separable Gaussian in Q14 integers with reflect-101 borders -> int64 image (rounded)"""
    k, r = _gauss_kernel_q14(sigma)
    a = np.pad(img.astype(np.int64), ((0, 0), (r, r)), mode="reflect")
    acc = np.zeros(img.shape, np.int64)
    for i, w in enumerate(k):
        acc += w * a[:, i:i + img.shape[1]]
    a = np.pad((acc + 8192) >> 14, ((r, r), (0, 0)), mode="reflect")
    acc = np.zeros(img.shape, np.int64)
    for i, w in enumerate(k):
        acc += w * a[i:i + img.shape[0], :]
    return (acc + 8192) >> 14


def _homography(src, dst):
    """3x3 H with H*src_i ~ dst_i (4 points), float64"""
    A = []
    b = []
    for (x, y), (u, v) in zip(src, dst):
        A.append([x, y, 1, 0, 0, 0, -u * x, -u * y]); b.append(u)
        A.append([0, 0, 0, x, y, 1, -v * x, -v * y]); b.append(v)
    h = np.linalg.solve(np.array(A, np.float64), np.array(b, np.float64))
    return np.append(h, 1.0).reshape(3, 3)


def _draw_marker(img, cells, corners):
    """paint the marker (with a one-cell white quiet zone) whose BLACK-BORDER outer corners map to `corners`
    (4x2, clockwise from the marker's top-left)"""
    n = cells.shape[0]
    full = np.pad(cells, 1, constant_values=1)          # quiet zone
    m = n + 2
    # marker coordinates: black border spans [0,n]; quiet zone spans [-1,n+1]
    src = np.array([[0, 0], [n, 0], [n, n], [0, n]], np.float64)
    H = _homography(src, corners)
    Hi = np.linalg.inv(H)
    q = np.array([[-1, -1, 1], [n + 1, -1, 1], [n + 1, n + 1, 1], [-1, n + 1, 1]], np.float64) @ H.T
    q = q[:, :2] / q[:, 2:3]
    x0 = max(int(np.floor(q[:, 0].min())), 0); x1 = min(int(np.ceil(q[:, 0].max())) + 1, img.shape[1])
    y0 = max(int(np.floor(q[:, 1].min())), 0); y1 = min(int(np.ceil(q[:, 1].max())) + 1, img.shape[0])
    if x1 <= x0 or y1 <= y0:
        return
    # 3x3 supersampling for clean edges
    sub = (np.arange(3) + 0.5) / 3.0 - 0.5
    acc = np.zeros((y1 - y0, x1 - x0), np.float64)
    cov = np.zeros((y1 - y0, x1 - x0), np.float64)
    ys, xs = np.mgrid[y0:y1, x0:x1].astype(np.float64)
    for dy in sub:
        for dx in sub:
            X = xs + dx; Y = ys + dy
            W = Hi[2, 0] * X + Hi[2, 1] * Y + Hi[2, 2]
            U = (Hi[0, 0] * X + Hi[0, 1] * Y + Hi[0, 2]) / W
            V = (Hi[1, 0] * X + Hi[1, 1] * Y + Hi[1, 2]) / W
            cu = np.floor(U).astype(np.int64) + 1
            cv = np.floor(V).astype(np.int64) + 1
            inside = (cu >= 0) & (cu < m) & (cv >= 0) & (cv < m)
            val = full[np.clip(cv, 0, m - 1), np.clip(cu, 0, m - 1)].astype(np.float64)
            acc += np.where(inside, val, 0.0)
            cov += inside
    region = img[y0:y1, x0:x1].astype(np.float64)
    white, black = 235.0, 20.0
    col = black + (white - black) * np.where(cov > 0, acc / np.maximum(cov, 1), 0.0)
    a = cov / 9.0
    img[y0:y1, x0:x1] = np.clip(np.floor(region * (1 - a) + col * a + 0.5), 0, 255).astype(np.uint8)


def make_frame(index, w=640, h=480, markers=0, dict_name="ARUCO_MIP_25h7", shift=(0.0, 0.0), return_truth=False):
    """one synthetic frame; markers = how many fiducials to plant (0 = extract-only workloads)"""
    rng = np.random.Generator(np.random.PCG64(20260000 + int(index)))
    s = w / 640.0
    bg = rng.integers(0, 256, size=(h, w), dtype=np.int64)
    bg = _blur_int(bg, 1.5)
    lo, hi = int(bg.min()), int(bg.max())
    img = (30 + (bg - lo) * 195 // max(hi - lo, 1)).astype(np.uint8)
    nshape = 150
    for _ in range(nshape):
        kind = int(rng.integers(0, 2))
        cx = int(rng.integers(0, w)); cy = int(rng.integers(0, h))
        sw = int(rng.integers(int(8 * s), int(60 * s) + 1)); sh = int(rng.integers(int(8 * s), int(60 * s) + 1))
        g = int(rng.integers(0, 256))
        x0, x1 = max(cx - sw // 2, 0), min(cx + (sw + 1) // 2, w)
        y0, y1 = max(cy - sh // 2, 0), min(cy + (sh + 1) // 2, h)
        if x1 <= x0 or y1 <= y0:
            continue
        if kind == 0:
            img[y0:y1, x0:x1] = g
        else:
            ys, xs = np.mgrid[y0:y1, x0:x1]
            # right triangle with a random orientation: integer half-plane test
            fx = int(rng.integers(0, 2)); fy = int(rng.integers(0, 2))
            u = (xs - x0) if fx == 0 else (x1 - 1 - xs)
            v = (ys - y0) if fy == 0 else (y1 - 1 - ys)
            mask = u * (y1 - y0) + v * (x1 - x0) <= (x1 - x0) * (y1 - y0)
            region = img[y0:y1, x0:x1]
            region[mask] = g
    truth = []
    if markers > 0:
        nbits, _, codes = dictionaries()[dict_name]
        ids = rng.permutation(min(len(codes), 100 if dict_name != "ARUCO" else len(codes)))[:markers]
        gx, gy = 5, 4
        slots = rng.permutation(gx * gy)[:markers]
        cw, ch = w / gx, h / gy
        for mid, slot in zip(ids, slots):
            side = float(rng.uniform(50 * s, min(76 * s, 0.64 * min(cw, ch))))
            ang = float(rng.uniform(-np.pi, np.pi))
            cxm = (slot % gx + 0.5) * cw + float(rng.uniform(-0.08, 0.08)) * cw
            cym = (slot // gx + 0.5) * ch + float(rng.uniform(-0.08, 0.08)) * ch
            base = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float64) * (side / 2)
            R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
            c = base @ R.T + rng.uniform(-0.08, 0.08, size=(4, 2)) * side
            c[:, 0] += cxm; c[:, 1] += cym
            # keep everything (incl. quiet zone) >= 12 px from the frame border
            margin = 12 * s + side / 5
            dx = max(margin - c[:, 0].min(), 0) - max(c[:, 0].max() - (w - 1 - margin), 0)
            dy = max(margin - c[:, 1].min(), 0) - max(c[:, 1].max() - (h - 1 - margin), 0)
            c[:, 0] += dx; c[:, 1] += dy
            _draw_marker(img, marker_cells(dict_name, int(mid)), c)
            truth.append((int(mid), c.copy()))
    if shift != (0.0, 0.0):
        img = np.roll(img, (int(shift[1]), int(shift[0])), axis=(0, 1))
    out = _blur_int(img, 0.7)
    noise = np.floor(rng.normal(0.0, 2.0, size=(h, w)) + 0.5).astype(np.int64)
    out = np.clip(out + noise, 0, 255).astype(np.uint8)
    if return_truth:
        return out, truth
    return out


def make_batch(n, w=640, h=480, markers=0, dict_name="ARUCO_MIP_25h7", first=0):
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        out[i] = make_frame(first + i, w, h, markers, dict_name)
    return out


def make_view(scene, index, max_rot_deg=3.0, max_shift=8.0, noise_sigma=2.0, rot_deg=None, shift=None):
    """A perturbed view of `scene` (u8 [h, w]): in-plane rotation about the image centre, translation, fresh sensor noise - the camera moved a little.
    Bilinear resampling in float64 with reflected borders; view `index` uses Generator(PCG64(20270000 + index)) unless rot_deg / shift are given."""
    rng = np.random.Generator(np.random.PCG64(20270000 + int(index)))
    h, w = scene.shape
    a = np.deg2rad(float(rng.uniform(-max_rot_deg, max_rot_deg)) if rot_deg is None else float(rot_deg))
    tx, ty = (rng.uniform(-max_shift, max_shift, size=2) if shift is None else shift)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    ca, sa = np.cos(a), np.sin(a)
    X = ca * (xs - cx - tx) + sa * (ys - cy - ty) + cx          # inverse map: output pixel -> scene position
    Y = -sa * (xs - cx - tx) + ca * (ys - cy - ty) + cy
    x0 = np.floor(X); y0 = np.floor(Y)
    fx = X - x0; fy = Y - y0

    def refl(i, n):
        i = np.abs(i.astype(np.int64))
        i = np.where(i >= n, 2 * (n - 1) - i, i)
        return np.clip(i, 0, n - 1)
    xa, xb, ya, yb = refl(x0, w), refl(x0 + 1, w), refl(y0, h), refl(y0 + 1, h)
    src = scene.astype(np.float64)
    v = (src[ya, xa] * (1 - fx) + src[ya, xb] * fx) * (1 - fy) + (src[yb, xa] * (1 - fx) + src[yb, xb] * fx) * fy
    if noise_sigma > 0:
        v = v + np.floor(rng.normal(0.0, noise_sigma, size=(h, w)) + 0.5)
    return np.clip(np.floor(v + 0.5), 0, 255).astype(np.uint8)
