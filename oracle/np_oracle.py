"""np_oracle.py — second, independent restatement of the hot path in numpy float32 (TEST INFRASTRUCTURE).

Written from the reference's shader text (reference src/shaders.metal:51-67, 87-95, 115-156, 159-186, 245-368), not
from mm_oracle.cpp, and vectorised over paths instead of scalar, so that an error of transcription in one of the two
shows up as a disagreement (tests/test_oracle.py compares them path by path: first-hit ids, segment counts, mirror
hits and radiance bit for bit).  numpy float32 add/sub/mul/div/sqrt are single IEEE-754 round-to-nearest operations,
the same canonical arithmetic as SURVEY §8 a-0.

PARITY UNPINNED BY THE REFERENCE (it has no tests or fixtures for this path); see mm_oracle.cpp's header.
"""
import numpy as np

F = np.float32
U = np.uint32


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _length(a):
    return np.sqrt(_dot(a, a))


def _normalize(a):
    l = _length(a)
    return [a[0] / l, a[1] / l, a[2] / l]


def _sign(x):
    return np.where(x > 0, F(1), np.where(x < 0, F(-1), F(0))).astype(F)


def _f2u_sat(f):
    """float -> uint: truncate toward zero, saturate, NaN -> 0 (Metal / CUDA conversion; SURVEY a-0)."""
    f64 = f.astype(np.float64)
    f64 = np.where(np.isnan(f64), 0.0, f64)
    return np.clip(np.trunc(f64), 0.0, 4294967295.0).astype(np.uint64).astype(U)


def _random(state):
    """shaders.metal:181-186 on a uint32 array; returns (new_state, float32 in [0,1])."""
    state = state * U(747796405) + U(291336453)
    shift = (state >> U(28)) + U(4)
    result = ((state >> shift) ^ state) * U(277803737)
    result = (result >> U(22)) ^ result
    return state, result.astype(F) / F(4294967296.0)


RCP_SLAB = False   # set by render() from params.flags & 64 (MM_FLAG_RCP_SLAB): t = (b - o) * (1/d) instead of (b - o) / d


def _aabb(o, d, t, bmin, bmax):
    """shaders.metal:87-95, vectorised."""
    if RCP_SLAB:
        i = [F(1.0) / d[0], F(1.0) / d[1], F(1.0) / d[2]]
        tx1 = (bmin[0] - o[0]) * i[0]; tx2 = (bmax[0] - o[0]) * i[0]
        tmin = np.fmin(tx1, tx2); tmax = np.fmax(tx1, tx2)
        ty1 = (bmin[1] - o[1]) * i[1]; ty2 = (bmax[1] - o[1]) * i[1]
        tmin = np.fmax(tmin, np.fmin(ty1, ty2)); tmax = np.fmin(tmax, np.fmax(ty1, ty2))
        tz1 = (bmin[2] - o[2]) * i[2]; tz2 = (bmax[2] - o[2]) * i[2]
        tmin = np.fmax(tmin, np.fmin(tz1, tz2)); tmax = np.fmin(tmax, np.fmax(tz1, tz2))
        ok = (tmax >= tmin) & (tmin < t) & (tmax > 0)
        return np.where(ok, tmin, F(1e30)).astype(F)
    tx1 = (bmin[0] - o[0]) / d[0]; tx2 = (bmax[0] - o[0]) / d[0]
    tmin = np.fmin(tx1, tx2); tmax = np.fmax(tx1, tx2)
    ty1 = (bmin[1] - o[1]) / d[1]; ty2 = (bmax[1] - o[1]) / d[1]
    tmin = np.fmax(tmin, np.fmin(ty1, ty2)); tmax = np.fmin(tmax, np.fmax(ty1, ty2))
    tz1 = (bmin[2] - o[2]) / d[2]; tz2 = (bmax[2] - o[2]) / d[2]
    tmin = np.fmax(tmin, np.fmin(tz1, tz2)); tmax = np.fmin(tmax, np.fmax(tz1, tz2))
    ok = (tmax >= tmin) & (tmin < t) & (tmax > 0)
    return np.where(ok, tmin, F(1e30)).astype(F)


def _rect(o, d, t, index, rect, pi):
    """shaders.metal:51-67, vectorised; rect = dict of per-lane (origin, v, u) component lists."""
    n = _normalize(_cross(rect["v"], rect["u"]))
    nc = _dot(d, n)
    a = _dot([rect["o"][k] - o[k] for k in range(3)], n) / nc
    p = [o[k] + a * d[k] for k in range(3)]
    rv = [p[k] - rect["o"][k] for k in range(3)]
    lv, lu = _length(rect["v"]), _length(rect["u"])
    d1 = _dot(rv, rect["v"]) / lv
    d2 = _dot(rv, rect["u"]) / lu
    hit = (0 <= d1) & (d1 <= lv) & (0 <= d2) & (d2 <= lu) & (nc != 0) & (a > F(0.1)) & (a < t)
    return np.where(hit, a, t).astype(F), np.where(hit, pi, index).astype(U)


def _traverse(sc, o, d, t, index, counters):
    """shaders.metal:115-156 for a batch of rays (lists of float32 arrays); returns (t, index)."""
    n = len(t)
    nodes = sc.nodes
    bmin_all, bmax_all = nodes["aabb_min"], nodes["aabb_max"]
    lf_all, tc_all = nodes["left_first"], nodes["tri_count"]
    planes, indices = sc.planes, sc.indices
    node = np.zeros(n, dtype=np.int64)
    stack = np.zeros((n, 50), dtype=np.int64)
    head = np.zeros(n, dtype=np.int64)
    alive = np.ones(n, dtype=bool)
    counters["rays"] += n
    while alive.any():
        ids = np.nonzero(alive)[0]
        cnt = tc_all[node[ids]]
        leaf = ids[cnt > 0]
        inner = ids[cnt == 0]
        if len(leaf):
            counters["leaf_visits"] += len(leaf)
            lf, tc = lf_all[node[leaf]].astype(np.int64), tc_all[node[leaf]].astype(np.int64)
            for i in range(int(tc.max())):
                m = tc > i
                L = leaf[m]
                pi = indices[lf[m] + i]
                pl = planes[pi]
                rect = {"o": [pl["origin"][:, k] for k in range(3)], "v": [pl["v"][:, k] for k in range(3)],
                        "u": [pl["u"][:, k] for k in range(3)]}
                nt, ni = _rect([c[L] for c in o], [c[L] for c in d], t[L], index[L], rect, pi.astype(U))
                t[L] = nt
                index[L] = ni
                counters["rect_tests"] += len(L)
            empty = head[leaf] == 0
            alive[leaf[empty]] = False
            pop = leaf[~empty]
            head[pop] -= 1
            node[pop] = stack[pop, head[pop]]
        if len(inner):
            counters["inner_visits"] += len(inner)
            left = lf_all[node[inner]].astype(np.int64)
            right = left + 1
            oi, di, ti = [c[inner] for c in o], [c[inner] for c in d], t[inner]
            d1 = _aabb(oi, di, ti, [bmin_all[left, k] for k in range(3)], [bmax_all[left, k] for k in range(3)])
            d2 = _aabb(oi, di, ti, [bmin_all[right, k] for k in range(3)], [bmax_all[right, k] for k in range(3)])
            swap = d1 > d2
            d1s = np.where(swap, d2, d1); d2s = np.where(swap, d1, d2)
            near = np.where(swap, right, left); far = np.where(swap, left, right)
            miss = d1s == F(1e30)
            # near child missed: pop or finish
            mi = inner[miss]
            empty = head[mi] == 0
            alive[mi[empty]] = False
            pop = mi[~empty]
            head[pop] -= 1
            node[pop] = stack[pop, head[pop]]
            # descend
            hi = inner[~miss]
            node[hi] = near[~miss]
            push = ~miss & (d2s != F(1e30))
            pu = inner[push]
            stack[pu, head[pu]] = far[push]
            head[pu] += 1
            if len(pu):
                counters["max_stack"] = max(counters["max_stack"], int(head[pu].max()))
    return t, index


def render(scene, noise, uniform, params, chunks):
    """Full-grid render (group_first/step/count are honoured).  Returns (image, counters, debug)."""
    global RCP_SLAB
    RCP_SLAB = bool(params.flags & 64)
    try:
        with np.errstate(all="ignore"):
            return _render(scene, noise, uniform, params, chunks)
    finally:
        RCP_SLAB = False


def _render(sc, noise, U_, P, chunks):
    W, H = F(U_.view_width), F(U_.view_height)
    chunk = int(U_.chunk_width)
    ppc = chunk * chunk
    spp = int(P.spp)
    T = ppc * spp
    assert T <= 32 or T % 32 == 0, "virtual threadgroup is (32, T/32)"
    dimx = min(32, T)
    dimy = T // dimx
    n_groups = P.grid_x * P.grid_y
    first, step, count = P.group_first, (P.group_step or 1), P.group_count
    if count == 0:
        first, step, count = 0, 1, n_groups
    k = np.repeat(np.arange(count, dtype=np.int64), T)
    flat = np.tile(np.arange(T, dtype=np.int64), count)
    g = first + k * step
    tgx, tgy = g % P.grid_x, g // P.grid_x
    gx, gy = flat % dimx, flat // dimx
    ch = chunks[tgx + tgy * P.grid_x]                                  # :266-267 (row stride = grid width)
    pn = flat // (T // ppc)                                            # :272
    px = ch["x"].astype(np.int64) + pn // chunk                        # :274-275
    py = ch["y"].astype(np.int64) + pn % chunk                         # :273,275
    texx = (tgx * dimx + gx).astype(U)
    texy = (tgy * dimy + gy).astype(U)
    n = len(flat)

    cam = U_.cam
    center = [F(cam.camera_center.x), F(cam.camera_center.y), F(cam.camera_center.z)]
    vpx, vpy, focal = F(cam.viewport.x), F(cam.viewport.y), F(cam.focal_length)
    pnx = px.astype(U).astype(F) / W                                   # :281
    pny = py.astype(U).astype(F) / H
    corner = [center[0] - vpx / F(2.0), center[1] - vpy / F(2.0), center[2] - (-focal)]   # :282
    rd = _normalize([(corner[0] + pnx * vpx) - center[0], (corner[1] + pny * vpy) - center[1],
                     np.full(n, (corner[2] + F(0.0)) - center[2], dtype=F)])               # :283
    q = [F(cam.rotation.x), F(cam.rotation.y), F(cam.rotation.z), F(cam.rotation.w)]

    def quat_dot(q1, q2):                                              # :163-167
        s = q1[3] * q2[3] - _dot(q1[:3], q2[:3])
        c = _cross(q1[:3], q2[:3])
        v = [(c[i] + q1[3] * q2[i]) + q2[3] * q1[i] for i in range(3)]
        return v + [s]

    zero = np.zeros(n, dtype=F)
    inv = [-q[0], -q[1], -q[2], q[3]]
    r = quat_dot(quat_dot(inv, rd + [zero]), q)                        # :169-172
    rd = [np.asarray(r[0], dtype=F), np.asarray(r[1], dtype=F), np.asarray(r[2], dtype=F)]

    # :288-291 sampler(address::repeat, filter::nearest), normalised coordinates
    nh, nw = noise.shape[:2]
    u, v = gx.astype(F), gy.astype(F)
    fu, fv = u - np.floor(u), v - np.floor(v)
    ix = np.clip(np.floor(fu * F(nw)).astype(np.int64), 0, nw - 1)
    iy = np.clip(np.floor(fv * F(nh)).astype(np.int64), 0, nh - 1)
    tex = noise[iy, ix].astype(F) / F(255.0)
    seed_f = (((tex[:, 0] + tex[:, 1]) + (texx * U(15823)).astype(F)) + (texy * U(9737333)).astype(F)) + F(U_.time)   # :298
    state = _f2u_sat(seed_f)

    o = [np.full(n, center[i], dtype=F) for i in range(3)]             # :302
    state, r1 = _random(state)
    state, r2 = _random(state)
    d = [rd[0] + ((r1 - F(0.5)) * F(2.0)) * F(0.001), rd[1] + ((r2 - F(0.5)) * F(2.0)) * F(0.001), rd[2] + F(0.0) * F(0.001)]   # :303
    d = [np.asarray(c, dtype=F) for c in d]
    t = np.full(n, F(1e30), dtype=F)
    index = np.full(n, 0xFFFFFFFF, dtype=U)
    color = [np.ones(n, dtype=F) for _ in range(3)]
    light = [np.zeros(n, dtype=F) for _ in range(3)]
    mirror_hits = np.zeros(n, dtype=np.int64)
    segments = np.zeros(n, dtype=U)
    first_hit = np.full(n, 0xFFFFFFFF, dtype=U)
    running = np.ones(n, dtype=bool)
    counters = {"paths": n, "rays": 0, "inner_visits": 0, "leaf_visits": 0, "rect_tests": 0, "hits": 0, "max_stack": 0}
    planes, materials, emissions = sc.planes, sc.materials, sc.emissions
    bl, ml = int(P.bounce_limit), int(P.mirror_limit)
    it = 0
    while True:
        act = running & (it < bl + mirror_hits)                        # :306
        running &= act
        A = np.nonzero(act)[0]
        if len(A) == 0:
            break
        ta, ia = _traverse(sc, [c[A] for c in o], [c[A] for c in d], t[A].copy(), index[A].copy(), counters)   # :307
        t[A], index[A] = ta, ia
        segments[A] += U(1)
        hit = ta < F(1e30)                                             # :308
        running[A[~hit]] = False                                       # :336-339
        Hh = A[hit]
        counters["hits"] += len(Hh)
        if it == 0:
            first_hit[Hh] = index[Hh]
        if len(Hh):
            pl = planes[index[Hh]]
            nrm = _normalize(_cross([pl["v"][:, kk] for kk in range(3)], [pl["u"][:, kk] for kk in range(3)]))   # :309
            dh = [c[Hh] for c in d]
            side = -_sign(_dot(dh, nrm))                               # :310
            diffuse = (materials[index[Hh]] == 0) | (side == F(-1.0))  # :311
            # --- diffuse branch :312-323
            D = Hh[diffuse]
            if len(D):
                e = emissions[index[D]]
                nd = [c[diffuse] for c in nrm]
                sd = side[diffuse]
                for kk in range(3):
                    emitted = e[:, kk] * e[:, 3]
                    light[kk][D] = light[kk][D] + emitted * color[kk][D]
                    color[kk][D] = color[kk][D] * pl["color"][diffuse, kk]
                st = state[D]
                rnd = [np.zeros(len(D), dtype=F) for _ in range(3)]
                need = np.ones(len(D), dtype=bool)
                while need.any():
                    idx = np.nonzero(need)[0]
                    s = st[idx]
                    comps = []
                    for _ in range(3):
                        s, rr = _random(s)
                        comps.append((rr - F(0.5)) * F(2.0))
                    st[idx] = s
                    for kk in range(3):
                        rnd[kk][idx] = comps[kk]
                    need[idx] = _length(comps) > F(1.0)
                state[D] = st
                rnd = _normalize(rnd)                                  # :319
                for kk in range(3):
                    o[kk][D] = o[kk][D] + d[kk][D] * t[D]              # :320
                nd2 = _normalize([rnd[kk] + nd[kk] * sd for kk in range(3)])   # :321
                for kk in range(3):
                    d[kk][D] = nd2[kk]
                t[D] = F(1e30)
            # --- mirror branch :325-334
            M = Hh[~diffuse]
            if len(M):
                mirror_hits[M] += 1
                cont = mirror_hits[M] < ml
                running[M[~cont]] = False
                Mc = M[cont]
                if len(Mc):
                    sel = np.nonzero(~diffuse)[0][cont]
                    nm = [c[sel] for c in nrm]
                    for kk in range(3):
                        light[kk][Mc] = light[kk][Mc] + planes["color"][index[Mc], kk] * F(0.005)   # :327
                    dm = [c[Mc] for c in d]
                    for kk in range(3):
                        o[kk][Mc] = o[kk][Mc] + dm[kk] * t[Mc]         # :328
                    two = F(2.0) * _dot(nm, dm)
                    refl = _normalize([dm[kk] - two * nm[kk] for kk in range(3)])   # :329 reflect(I,N) = I - 2*dot(N,I)*N
                    for kk in range(3):
                        d[kk][Mc] = refl[kk]
                    t[Mc] = F(1e30)
        it += 1

    radiance = np.stack(light, axis=1)
    tone = np.stack([np.sqrt(np.fmax(c, F(0.0))) for c in light], axis=1).reshape(count, T, 3)   # :344
    # :347-358 pairwise phases (only the strides < spp), then :360-366 serial octets and divide
    for stride in (1, 2, 4):
        if stride < spp:
            tone[:, ::2 * stride] = tone[:, ::2 * stride] + tone[:, stride::2 * stride]
    tone = tone.reshape(count, ppc, spp, 3)
    acc = tone[:, :, 0].copy()
    for i in range(1, spp // 8):
        acc = acc + tone[:, :, 8 * i]
    acc = acc / F(spp)
    image = np.zeros((int(H), int(W), 4), dtype=F)
    fx = px.reshape(count, ppc, spp)[:, :, 0]
    fy = py.reshape(count, ppc, spp)[:, :, 0]
    ok = (fx < int(W)) & (fy < int(H))
    image[fy[ok], fx[ok], :3] = acc[ok]
    image[fy[ok], fx[ok], 3] = F(1.0)
    debug = {"first_hit": first_hit, "segments": segments, "mirror_hits": mirror_hits.astype(U), "radiance": radiance}
    return image, counters, debug


def quant8(a):
    """Store into an RGBA8Unorm texture and read back: rte(clamp(v, 0, 1) * 255) / 255 in float32, NaN -> 0."""
    a = np.asarray(a, dtype=F)
    c = np.where(a > 0, a, F(0.0)).astype(F)
    c = np.where(c < 1, c, F(1.0)).astype(F)
    return (np.rint(c * F(255.0)).astype(F) / F(255.0)).astype(F)


def present_blur(img, rgba8=False):
    """fragment_shader (reference src/shaders.metal:214-225) as a race-free ping-pong pass over an [H,W,4] float32 image:
    c = img[p]; c += (img[p+(1,0)] + img[p-(1,0)]) / 2; c += (img[p+(0,1)] + img[p-(0,1)]) / 2; c /= 3; out = (c.rgb, 1).
    Reads outside the texture return 0."""
    img = np.asarray(img, dtype=F)
    H, W = img.shape[:2]
    pad = np.zeros((H + 2, W + 2, 4), dtype=F)
    pad[1:-1, 1:-1] = img
    c = pad[1:-1, 1:-1]
    r, l = pad[1:-1, 2:], pad[1:-1, :-2]
    d, u = pad[2:, 1:-1], pad[:-2, 1:-1]
    o = (c + (r + l) / F(2.0)) + (d + u) / F(2.0)
    o = o / F(3.0)
    o[..., 3] = F(1.0)
    return quant8(o) if rgba8 else o.astype(F)
