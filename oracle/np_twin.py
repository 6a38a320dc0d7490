"""Independent numpy twin of the C oracle -- TEST INFRASTRUCTURE (small cases only).

Written against SURVEY.md Appendix A, *not* as a transliteration of gridgcn_oracle.c: it uses
the sorted / CSR formulation (stable sort by voxel key, first-occurrence ranking, k-way merge,
stable sort by distance) that the sm_100a kernels use, so agreement between the two pins both
the literal restatement and the reformulation.  Parity status: see the header of
gridgcn_oracle.c (pinned against the reference's own kernel bodies; CAS unpinned).

Reference lines followed: gridify.cu:126-190,218-290; gridifyknn.cu:231-332;
gridify_up.cu:121-169,190-224 (paths relative to /root/reference/gridifyop/).
"""
import numpy as np

F = np.float32


def _voxelise(xyz, shift, voxel, grid):
    """A.1: per-axis fp32 add, fp32 divide, floor; reject outside [0, grid)."""
    t = (xyz.astype(F) + shift.astype(F)).astype(F)
    q = (t / voxel.astype(F)).astype(F)
    c = np.floor(q).astype(np.int64)
    ok = np.all((c >= 0) & (c.astype(F) < grid.astype(F)), axis=1)
    lin = c[:, 2] * (grid[0] * grid[1]) + c[:, 1] * grid[0] + c[:, 0]
    return c, lin, ok


class _Csr:
    """Per-cloud sorted voxel table: for every occupied voxel the ascending list of its points."""

    def __init__(self, data, npts, shift, voxel, grid):
        self.grid = grid
        pts = data[:npts]
        c, lin, ok = _voxelise(pts[:, :3], shift, voxel, grid)
        ids = np.nonzero(ok)[0]
        order = ids[np.argsort(lin[ids], kind="stable")]  # stable: ascending ids inside a voxel
        keys = lin[order]
        self.members = {}
        if len(order):
            cuts = np.nonzero(np.diff(keys))[0] + 1
            for seg in np.split(order, cuts):
                self.members[int(lin[seg[0]])] = seg
        # centre order = order of first occurrence of the voxel in point order
        first = sorted((int(seg[0]), v) for v, seg in self.members.items())
        self.center_voxels = [v for _, v in first]

    def get(self, d, h, w):
        g = self.grid
        if not (0 <= d < g[2] and 0 <= h < g[1] and 0 <= w < g[0]):
            return None
        return self.members.get(int(d * g[0] * g[1] + h * g[0] + w), np.empty(0, np.int64))


def _barycenter(data, seg):
    """A.2: sequential fp32 accumulation of (x*w, y*w, z*w, w) in ascending point order."""
    acc = np.zeros(4, F)
    for i in seg:
        w = F(data[i, 3])
        acc[0] = F(acc[0] + F(F(data[i, 0]) * w))
        acc[1] = F(acc[1] + F(F(data[i, 1]) * w))
        acc[2] = F(acc[2] + F(F(data[i, 2]) * w))
        acc[3] = F(acc[3] + w)
    return acc


def _outputs(B, O, P):
    return (np.zeros((B, O, P), np.int32), np.zeros((B, O, P), F), np.ones((B, O, 4), F),
            np.zeros((B, O), F), np.zeros((B, 1), np.int32))


def _decode(v, grid):
    c2 = v // (grid[0] * grid[1])
    c1 = (v - c2 * grid[0] * grid[1]) // grid[0]
    c0 = v - c2 * grid[0] * grid[1] - c1 * grid[0]
    return int(c2), int(c1), int(c0)


def _center_xyz(data, csr, v, loc, cent_row):
    if loc == 1:
        acc = _barycenter(data, csr.members[v])
        cent_row[0] = F(acc[0] / acc[3])
        cent_row[1] = F(acc[1] / acc[3])
        cent_row[2] = F(acc[2] / acc[3])


def gridify(data, npts, *, max_p_grid, max_o_grid, kernel_size, loc, coord_shift, voxel_size,
            grid_size, select_centers=None):
    """``select_centers(csr) -> list of centre voxels`` overrides the keep-first centre choice
    (used by gridify_occaware below)."""
    data = np.asarray(data, F)
    B, N, _ = data.shape
    O, P, ks = max_o_grid, max_p_grid, kernel_size
    shift, voxel, grid = (np.asarray(coord_shift, F), np.asarray(voxel_size, F),
                          np.asarray(grid_size, np.int64))
    nebidx, nebmsk, cent, centmsk, centnum = _outputs(B, O, P)
    r = (ks - 1) // 2
    for b in range(B):
        csr = _Csr(data[b], int(np.asarray(npts).reshape(-1)[b]), shift, voxel, grid)
        if select_centers is not None:
            csr.center_voxels = select_centers(csr)
        nc = min(len(csr.center_voxels), O)
        centnum[b, 0] = nc
        centmsk[b, :nc] = 1.0
        for o in range(nc):
            v = csr.center_voxels[o]
            c2, c1, c0 = _decode(v, grid)
            ids = []
            for t in range(ks ** 3):  # raster order d -> h -> w, A.3
                seg = csr.get(t // (ks * ks) - r + c2, (t % (ks * ks)) // ks - r + c1,
                              t % ks - r + c0)
                if seg is not None:
                    ids.extend(seg[:P].tolist())
            ids = ids[:P]  # keep-first
            n = len(ids)
            nebidx[b, o, :n] = ids
            nebidx[b, o, n:] = ids[0]
            nebmsk[b, o, :n] = 1.0
            wsum = F(0)
            for i in ids:
                wsum = F(wsum + F(int(data[b, i, 3])))
            cent[b, o, 3] = wsum
            _center_xyz(data[b], csr, v, loc, cent[b, o])
    return nebidx, nebmsk, cent, centmsk, centnum


def _dist2(u, p, fma):
    dx, dy, dz = F(u[0] - p[0]), F(u[1] - p[1]), F(u[2] - p[2])
    if fma:  # fma(dz,dz, fma(dy,dy, dx*dx)) evaluated exactly in float64 then rounded once each
        a = F(dx * dx)
        b = F(np.float64(dy) * np.float64(dy) + np.float64(a))
        return F(np.float64(dz) * np.float64(dz) + np.float64(b))
    return F(F(F(dx * dx) + F(dy * dy)) + F(dz * dz))


def gridify_knn(data, npts, *, max_p_grid, max_o_grid, kernel_size, loc, coord_shift, voxel_size,
                grid_size, dist_fma=False):
    data = np.asarray(data, F)
    B, N, _ = data.shape
    O, P, ks = max_o_grid, max_p_grid, kernel_size
    shift, voxel, grid = (np.asarray(coord_shift, F), np.asarray(voxel_size, F),
                          np.asarray(grid_size, np.int64))
    nebidx, nebmsk, cent, centmsk, centnum = _outputs(B, O, P)
    for b in range(B):
        csr = _Csr(data[b], int(np.asarray(npts).reshape(-1)[b]), shift, voxel, grid)
        nc = min(len(csr.center_voxels), O)
        centnum[b, 0] = nc
        centmsk[b, :nc] = 1.0
        for o in range(nc):
            v = csr.center_voxels[o]
            c2, c1, c0 = _decode(v, grid)
            u = [F((np.float64(c) + 0.5) * np.float64(vs)) for c, vs in zip((c0, c1, c2), voxel)]
            cand = []  # arrival order
            for layer in range((ks + 1) // 2):
                for w in range(-layer, layer + 1):
                    for h in range(-layer, layer + 1):
                        for d in range(-layer, layer + 1):
                            if max(abs(w), abs(h), abs(d)) != layer:
                                continue
                            seg = csr.get(d + c2, h + c1, w + c0)
                            if seg is not None:
                                cand.extend(seg[:P].tolist())
                if len(cand) >= P:
                    break
            dst = np.array([_dist2(u, data[b, i, :3], dist_fma) for i in cand], F)
            order = np.argsort(dst, kind="stable")[:P]  # strict-< insertion == stable sort
            ids = [cand[j] for j in order]
            n = len(ids)
            nebidx[b, o, :n] = ids
            nebidx[b, o, n:] = ids[0]
            nebmsk[b, o, :] = 1.0  # gridifyknn.cu:312: mask 1 for all P slots
            wsum = F(0)
            for i in ids:
                wsum = F(wsum + F(int(data[b, i, 3])))
            cent[b, o, 3] = wsum
            _center_xyz(data[b], csr, v, loc, cent[b, o])
    return nebidx, nebmsk, cent, centmsk, centnum


def gridify_up(downdata, updata, downnum, upnum, *, max_p_grid, max_o_grid, kernel_size,
               coord_shift, voxel_size, grid_size):
    """A.5 via the merge formulation: the bucket of voxel u holds, in ascending id order, the down
    points of every in-grid voxel of u's kernel neighbourhood (each point is splatted into u at
    most once), truncated to P."""
    downdata, updata = np.asarray(downdata, F), np.asarray(updata, F)
    B = downdata.shape[0]
    O, P, ks = max_o_grid, max_p_grid, kernel_size
    shift, voxel, grid = (np.asarray(coord_shift, F), np.asarray(voxel_size, F),
                          np.asarray(grid_size, np.int64))
    nebidx, nebmsk = np.zeros((B, O, P), np.int32), np.zeros((B, O, P), F)
    r = (ks - 1) // 2
    for b in range(B):
        csr = _Csr(downdata[b], int(np.asarray(downnum).reshape(-1)[b]), shift, voxel, grid)
        nu = min(int(np.asarray(upnum).reshape(-1)[b]), O)
        c, lin, ok = _voxelise(updata[b, :nu, :3], shift, voxel, grid)
        for o in range(nu):
            if not ok[o]:
                continue
            ids = []
            for t in range(ks ** 3):
                seg = csr.get(c[o, 2] + t // (ks * ks) - r, c[o, 1] + (t % (ks * ks)) // ks - r,
                              c[o, 0] + t % ks - r)
                if seg is not None:
                    ids.extend(seg.tolist())
            ids = sorted(ids)[:P]
            n = len(ids)
            if n:
                nebidx[b, o, :n] = ids
                nebidx[b, o, n:] = ids[0]
                nebmsk[b, o, :n] = 1.0
    return nebidx, nebmsk


def knn(unknown, known, downnum, upnum, *, k, radius=None, dist_fma=False):
    unknown, known = np.asarray(unknown, F), np.asarray(known, F)
    B, n, _ = unknown.shape
    idx = np.zeros((B, n, k), np.int32)
    for b in range(B):
        dn, un = int(np.asarray(downnum).reshape(-1)[b]), int(np.asarray(upnum).reshape(-1)[b])
        for q in range(min(un, n)):
            d = np.array([_dist2(unknown[b, q], known[b, j], dist_fma) for j in range(dn)], F)
            keep = np.arange(dn)
            if radius is not None:
                keep = keep[~(d > F(F(radius) * F(radius)))]
            order = keep[np.argsort(d[keep], kind="stable")][:k]
            idx[b, q, :] = -1 if radius is not None else 0
            idx[b, q, :len(order)] = order
    return idx


# ---------------------------------------------------------------------------------------------
# Coverage-Aware Sampling twin (Gridify_occaware; binary-only in the reference, see the CAS block
# comment of gridgcn_oracle.c for what the SASS of additional.so shows).  Formulated with a
# dictionary of coverage counts and Python-integer XORWOW, independently of the C arrays.
# ---------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def _xorwow_first_uniform(seed):
    """curand_init(seed, 0, 0) + one curand_uniform (curand_kernel.h:772-798,863-874;
    curand_uniform.h:69-72 contracted to one FMA on the device)."""
    s0 = (seed & _M32) ^ 0xaad26b49
    s1 = ((seed >> 32) & _M32) ^ 0xf7dcefdd
    t0 = (1099087573 * s0) & _M32
    t1 = (2591861531 * s1) & _M32
    d = (6615241 + t1 + t0) & _M32
    v = [(123456789 + t0) & _M32, 362436069 ^ t0, (521288629 + t1) & _M32, 88675123 ^ t1,
         (5783321 + t0) & _M32]
    t = v[0] ^ (v[0] >> 2)
    v4 = (v[4] ^ ((v[4] << 4) & _M32)) ^ (t ^ ((t << 1) & _M32))
    d = (d + 362437) & _M32
    x = (v4 + d) & _M32
    # single rounding of x * 2^-32 + 2^-33: exact in float64 (x < 2^32), then one rounding to f32
    return F(np.float64(F(x)) * np.float64(F(2.3283064e-10)) + np.float64(F(2.3283064e-10) / F(2.0)))


def cas_select(center_voxels, occupied, O, ks, grid, seed):
    """Returns the centre voxels after coverage-aware sampling: the first O occupied voxels (arrival
    order) are incumbents, every later one challenges a random incumbent once, in arrival order."""
    r = (ks - 1) // 2
    offs = [(t // (ks * ks) - r, (t % (ks * ks)) // ks - r, t % ks - r) for t in range(ks ** 3)]

    def nbrs(v):
        c2, c1, c0 = _decode(v, grid)
        out = []
        for od, oh, ow in offs:
            d, h, w = c2 + od, c1 + oh, c0 + ow
            if 0 <= d < grid[2] and 0 <= h < grid[1] and 0 <= w < grid[0]:
                out.append(int(d * grid[0] * grid[1] + h * grid[0] + w))
        return out

    slots = list(center_voxels[:O])
    cover = {}
    for v in slots:
        for n in nbrs(v):
            cover[n] = cover.get(n, 0) + 1

    def H(v, flip):
        h = F(0)
        for n in nbrs(v):
            if cover.get(n, 0) == flip:
                h = F(np.float64(h) + 0.7)
                if n in occupied:
                    h = F(np.float64(h) + 0.3)
        return h

    for i in range(O, len(center_voxels)):
        chal = center_voxels[i]
        u = _xorwow_first_uniform(seed + i)
        slot = int(np.ceil(F(F(O) * u)) - 1)
        inc = slots[slot]
        if H(chal, 0) > H(inc, 1):
            slots[slot] = chal
            for n in nbrs(inc):
                cover[n] -= 1
            for n in nbrs(chal):
                cover[n] = cover.get(n, 0) + 1
    return slots


def gridify_occaware(data, npts, *, max_p_grid, max_o_grid, kernel_size, loc, coord_shift, voxel_size,
                     grid_size, seed=0):
    grid = np.asarray(grid_size, np.int64)

    def select(csr):
        return cas_select(csr.center_voxels, set(csr.members.keys()), max_o_grid, kernel_size, grid, seed)

    return gridify(data, npts, max_p_grid=max_p_grid, max_o_grid=max_o_grid, kernel_size=kernel_size,
                   loc=loc, coord_shift=coord_shift, voxel_size=voxel_size, grid_size=grid_size,
                   select_centers=select)
