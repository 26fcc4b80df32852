"""The reference detector's stage sequence driven through the REAL cv2 primitives (dev container only).

Follows the de-obfuscated control flow of Thirdparty/aruco/aruco/markerdetector_impl.cpp:
adaptiveThreshold -> findContours -> approxPolyDP/isContourConvex -> prefilter -> pyramid -> warpPerspective ->
Otsu -> cell vote -> dictionary lookup -> sort/de-dup -> contour-line corner refinement (cv2.solve DECOMP_SVD).
Used by make_golden.py to produce tests/golden/aruco_pipeline.npz, the cross-check for oracle/aruco_oracle.cpp.
"""
import numpy as np
import cv2


def perimeter(c):
    s = 0
    for i in range(4):
        j = (i + 1) % 4
        dx = np.float32(c[i][0]) - np.float32(c[j][0]); dy = np.float32(c[i][1]) - np.float32(c[j][1])
        s += int(np.sqrt(np.float32(dx * dx + dy * dy)))
    return s


def get_area(c):
    c = c.astype(np.float32)
    v01 = c[1] - c[0]; v03 = c[3] - c[0]
    a1 = abs(np.float32(v01[0] * v03[1]) - np.float32(v01[1] * v03[0]))
    v21 = c[1] - c[2]; v23 = c[3] - c[2]
    a2 = abs(np.float32(v21[0] * v23[1]) - np.float32(v21[1] * v23[0]))
    return np.float32(np.float32(a2 + a1) / np.float32(2))


def detect(img, codes, nbits, solve=True):
    h, w = img.shape
    win = max(3, int(15 * np.float32(w) / 1920.))
    if win % 2 == 0:
        win += 1
    thres = cv2.adaptiveThreshold(img, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, win, 7)
    contours, _ = cv2.findContours(thres, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    cands = []
    for c in contours:
        if 70 < len(c):
            ap = cv2.approxPolyDP(c, len(c) * 0.05, True)
            if len(ap) == 4 and cv2.isContourConvex(ap):
                cands.append([ap.reshape(4, 2).astype(np.float32), c.reshape(-1, 2)])
    for cd in cands:
        c = cd[0].astype(np.float64)
        o = (c[1][0] - c[0][0]) * (c[2][1] - c[0][1]) - (c[1][1] - c[0][1]) * (c[2][0] - c[0][0])
        if o < 0:
            cd[0][[1, 3]] = cd[0][[3, 1]]
    near = []
    for i in range(len(cands)):
        for j in range(i + 1, len(cands)):
            d = [np.float32(np.sqrt(float(a[0] - b[0]) ** 2 + float(a[1] - b[1]) ** 2)) for a, b in zip(cands[i][0], cands[j][0])]
            if all(x < win for x in d):
                near.append((i, j))
    rm = [False] * len(cands)
    for i, j in near:
        if perimeter(cands[i][0]) > perimeter(cands[j][0]):
            rm[j] = True
        else:
            rm[i] = True
    bx, by = int(np.float32(0.015) * np.float32(w)), int(np.float32(0.015) * np.float32(h))
    for i, (c, _) in enumerate(cands):
        for p in c:
            if p[0] < bx or p[1] < by or p[0] > w - bx or p[1] > h - by:
                rm[i] = True
    kept = [cands[i] for i in range(len(cands)) if not rm[i]]
    nb = int(np.sqrt(nbits)); nsub = nb + 2; ws = 5 * nsub
    pyr = [img]
    cw, nl = w, 1
    while cw > 2 * ws:
        cw //= 2; nl += 1
    for l in range(1, nl):
        pyr.append(cv2.resize(pyr[-1], (pyr[-1].shape[1] // 2, pyr[-1].shape[0] // 2)))
    code_id = {}
    for i, c in enumerate(codes):
        code_id.setdefault(c, i)
    markers, patches = [], []
    for c, cont in kept:
        lvl = 0
        for p in range(1, len(pyr)):
            if get_area(c) / (4.0 ** p) >= np.float32(ws) ** 2:
                lvl = p
            else:
                break
        sc = (c * np.float32(np.float32(pyr[lvl].shape[1]) / np.float32(w))).astype(np.float32)
        dst = np.array([[0, 0], [ws - 1, 0], [ws - 1, ws - 1], [0, ws - 1]], np.float32)
        M = cv2.getPerspectiveTransform(sc, dst)
        patch = cv2.warpPerspective(pyr[lvl], M, (ws, ws), flags=cv2.INTER_LINEAR)
        patches.append(patch.copy())
        _, bw = cv2.threshold(patch, 125, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
        nz = np.zeros((nsub, nsub), int); tot = np.zeros((nsub, nsub), int)
        for y in range(ws):
            my = int(np.float32(nsub) * np.float32(y) / np.float32(ws))
            for x in range(ws):
                mx = int(np.float32(nsub) * np.float32(x) / np.float32(ws))
                nz[my, mx] += bw[y, x] > 125; tot[my, mx] += 1
        bits = (nz > tot // 2).astype(np.uint8)
        if bits[0].any() or bits[-1].any() or bits[:, 0].any() or bits[:, -1].any():
            continue
        inner = bits[1:-1, 1:-1].copy()
        ids = []
        for r in range(4):
            code = 0; b = 0
            for y in range(nb - 1, -1, -1):
                for x in range(nb - 1, -1, -1):
                    code |= int(inner[y, x]) << b; b += 1
            ids.append(code)
            t = np.zeros_like(inner)
            for i in range(nb):
                for j in range(nb):
                    t[i, j] = inner[nb - j - 1, i]
            inner = t
        if ids[0] == 0:
            continue
        for r in range(4):
            if ids[r] in code_id:
                markers.append([code_id[ids[r]], np.roll(c, -(4 - r) % 4, axis=0).copy() if False else np.array([c[(k + 4 - r) % 4] for k in range(4)], np.float32), cont])
                break
    order = sorted(range(len(markers)), key=lambda i: markers[i][0])      # stable
    markers = [markers[i] for i in order]
    rm = [False] * len(markers)
    for i in range(len(markers) - 1):
        j = i + 1
        while j < len(markers) and not rm[i]:
            if markers[i][0] == markers[j][0]:
                if perimeter(markers[i][1]) < perimeter(markers[j][1]):
                    rm[i] = True
                else:
                    rm[j] = True
            j += 1
    markers = [m for m, r in zip(markers, rm) if not r]
    pre = [m[1].copy() for m in markers]
    out = []
    for mid, c, cont in markers:
        n = len(cont)
        ci = [-1] * 4; md = [np.float32(3.4e38)] * 4
        for j in range(n):
            for k in range(4):
                dx = np.float32(cont[j][0]) - c[k][0]; dy = np.float32(cont[j][1]) - c[k][1]
                d = np.float32(np.float32(dx * dx) + np.float32(dy * dy))
                if d < md[k]:
                    ci[k] = j; md[k] = d
        if (ci[1] > ci[0]) and (ci[2] > ci[1] or ci[2] < ci[0]):
            inv = False
        elif ci[2] > ci[1] and ci[2] < ci[0]:
            inv = False
        else:
            inv = True
        inc = -1 if inv else 1
        lines = []
        for l in range(4):
            pts = []
            stop = ci[(l + 1) % 4]
            j = ci[l]
            while j != stop:
                if j == n and not inv:
                    j = 0
                elif j == 0 and inv:
                    j = n - 1
                pts.append(cont[j].astype(np.float32))
                if j == stop:
                    break
                j += inc
            pts = np.array(pts, np.float32)
            if pts[:, 0].max() - pts[:, 0].min() > pts[:, 1].max() - pts[:, 1].min():
                A = np.stack([pts[:, 0], np.ones(len(pts), np.float32)], 1); B = pts[:, 1:2].copy()
                _, X = cv2.solve(A, B, flags=cv2.DECOMP_SVD)
                lines.append((X[0, 0], np.float32(-1), X[1, 0]))
            else:
                A = np.stack([pts[:, 1], np.ones(len(pts), np.float32)], 1); B = pts[:, 0:1].copy()
                _, X = cv2.solve(A, B, flags=cv2.DECOMP_SVD)
                lines.append((np.float32(-1), X[0, 0], X[1, 0]))
        nc = []
        for i in range(4):
            l1, l2 = lines[(i - 1) % 4], lines[i]
            A = np.array([[l1[0], l1[1]], [l2[0], l2[1]]], np.float32); B = np.array([[-l1[2]], [-l2[2]]], np.float32)
            _, X = cv2.solve(A, B, flags=cv2.DECOMP_SVD)
            nc.append([X[0, 0], X[1, 0]])
        out.append((mid, np.array(nc, np.float32)))
    return dict(thres=thres, contours=[c.reshape(-1, 2) for c in contours], candidates=np.array([k[0] for k in kept], np.float32).reshape(-1, 4, 2),
                patches=np.array(patches, np.uint8), prerefine=np.array(pre, np.float32).reshape(-1, 4, 2), markers=out)
