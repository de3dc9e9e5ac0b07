"""TEST INFRASTRUCTURE ONLY -- a CPU (numpy, float64) restatement of the reference's FORWARD algorithm for the hot path,
restricted to what the golden cases C1 / C2 exercise: perspective look-at camera, Lambertian materials with constant
reflectance, area lights, max_bounces = 1 (next-event estimation + BSDF sampling with power-2 MIS), Sobol or PCG32
streams.  It is independent of redner_b200/csrc (no shared code) and brute-forces visibility (no BVH).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module, and only as the CHECKER.
The product path (redner_b200/) never imports it.

PARITY PINNING: this restatement is validated in tests/test_oracle_cpu.py against the golden vectors in tests/golden/,
which were produced by the compiled, unmodified reference (oracle/build_ref.sh + tests/golden/make_golden.py); the same
test re-derives those goldens bit-for-bit from oracle/_ref whenever it is available.  The backward pass / edge sampling
is NOT restated here: for those, the oracle is the compiled reference itself (oracle/_ref) and its golden vectors.

Every function cites the reference lines it follows (paths relative to the reference checkout).
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "redner_b200", "data", "sobol_joe_kuo_1024x52_u64.bin")
_sobol = None
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def sobol_matrices():
    """src/sobol.inc:32-35 (Joe-Kuo direction numbers, 1024 dims x 52 bits); table extracted by tools/extract_tables.py."""
    global _sobol
    if _sobol is None:
        _sobol = np.fromfile(_DATA, dtype=np.uint64).reshape(1024, 52)
    return _sobol


def hash64shift(key):
    """src/sobol_sampler.cpp:12-22"""
    key = np.asarray(key, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = (~key) + (key << np.uint64(21))
        key = key ^ (key >> np.uint64(24))
        key = (key + (key << np.uint64(3))) + (key << np.uint64(8))
        key = key ^ (key >> np.uint64(14))
        key = (key + (key << np.uint64(2))) + (key << np.uint64(4))
        key = key ^ (key >> np.uint64(28))
        key = key + (key << np.uint64(31))
    return key


def sobol_sample(index, dim, scramble):
    """src/sobol_sampler.cpp:61-76: one component of the scrambled Sobol sequence (vectorised over `scramble`)."""
    m = sobol_matrices()
    result = scramble & np.uint64((1 << 52) - 1)
    i = dim * 52
    flat = m.reshape(-1)
    while index:
        if index & 1:
            result = result ^ flat[i]
        index >>= 1
        i += 1
    return result.astype(np.float64) * (1.0 / float(1 << 52))


class SobolStream:
    """src/sobol_sampler.cpp:10-29 (per-pixel scramble), :97-100 (begin_sample): index = sample id, dimension advances
    with every next_* call."""

    def __init__(self, seed, num_pixels):
        idx = np.arange(num_pixels, dtype=np.uint64)
        self.scramble = hash64shift((np.uint64(seed) << np.uint64(32)) | idx)
        self.sample_id, self.dim = 0, 0

    def begin_sample(self, sample_id):
        self.sample_id, self.dim = sample_id, 0

    def next(self, n):
        out = np.stack([sobol_sample(self.sample_id, self.dim + i, self.scramble) for i in range(n)], axis=1)
        self.dim += n
        return out


class PCGStream:
    """src/pcg_sampler.cpp:8-50: one PCG32 stream per pixel, 32-bit draws mapped to double."""

    def __init__(self, seed, num_pixels):
        idx = np.arange(num_pixels, dtype=np.uint64)
        self.inc = ((idx + np.uint64(1)) << np.uint64(1)) | np.uint64(1)
        self.state = np.zeros(num_pixels, dtype=np.uint64)
        self._next32()
        with np.errstate(over="ignore"):
            self.state = self.state + (np.uint64(0x853c49e6748fea9b) + np.uint64(seed))
        self._next32()

    def begin_sample(self, sample_id):
        pass

    def _next32(self):
        old = self.state
        with np.errstate(over="ignore"):
            self.state = old * np.uint64(6364136223846793005) + (self.inc | np.uint64(1))
        xorshifted = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
        rot = (old >> np.uint64(59)).astype(np.uint32)
        return (xorshifted >> rot) | (xorshifted << ((np.uint32(0) - rot) & np.uint32(31)))

    def next(self, n):
        cols = []
        for _ in range(n):
            r = self._next32().astype(np.uint64)
            u = (r << np.uint64(20)) | np.uint64(0x3ff0000000000000)
            cols.append(u.view(np.float64) - 1.0)
        return np.stack(cols, axis=1)


def normalize(v):
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return np.where(n > 0, v / np.where(n > 0, n, 1), 0.0)


def look_at(pos, look, up):
    """src/transform.h:9-27"""
    d = normalize(look - pos)
    right = normalize(np.cross(d, normalize(up)))
    new_up = normalize(np.cross(right, d))
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, new_up, d, pos
    return m


def primary_rays(cam, samples):
    """src/camera.cpp:8-22 + src/camera.h:128-141 (perspective): pixel + jitter -> world-space ray."""
    h, w = cam["resolution"]
    idx = np.arange(h * w)
    px, py = idx % w, idx // w
    sx = (px + samples[:, 0]) / w
    sy = (py + samples[:, 1]) / h
    c2w = look_at(cam["position"], cam["look_at"], cam["up"])
    aspect = w / h
    pt = np.stack([(sx - 0.5) * 2.0, (sy - 0.5) * (-2.0) / aspect, np.ones_like(sx)], axis=1)
    d = normalize(pt @ cam["intrinsic_mat_inv"].T)
    wd = normalize(d @ c2w[:3, :3].T)
    org = np.broadcast_to(c2w[:3, 3] / c2w[3, 3], wd.shape)
    # the reference hands fp32 copies of the ray to Embree (src/scene.cpp:556-567); the hit DECISION is taken on those
    return org, wd


def intersect_all(tris, org, d, tmin, tmax):
    """Closest hit by brute force over all triangles with the Moeller-Trumbore solve of src/intersection.h:55-109
    (visibility in the reference comes from Embree, src/scene.cpp:546-592; away from silhouette pixels both agree).
    tris: [T, 3, 3].  Returns (triangle index or -1, t)."""
    n = org.shape[0]
    best_t = np.full(n, np.inf)
    best = np.full(n, -1)
    o32, d32 = org.astype(np.float32).astype(np.float64), d.astype(np.float32).astype(np.float64)
    for ti, (v0, v1, v2) in enumerate(tris):
        e1, e2 = v1 - v0, v2 - v0
        pvec = np.cross(d32, e2)
        div = pvec @ e1
        ok = np.abs(div) > 1e-30
        inv = np.where(ok, 1.0 / np.where(ok, div, 1.0), 0.0)
        s = o32 - v0
        u = np.einsum("ij,ij->i", s, pvec) * inv
        q = np.cross(s, e1)
        v = np.einsum("ij,ij->i", d32, q) * inv
        t = (q @ e2) * inv
        hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > tmin) & (t <= np.minimum(tmax, best_t))
        best_t = np.where(hit, t, best_t)
        best = np.where(hit, ti, best)
    return best, best_t


def occluded(tris, org, d, tmin, tmax):
    b, _ = intersect_all(tris, org, d, tmin, tmax)
    return b >= 0


def tri_normal(tri):
    return normalize(np.cross(tri[1] - tri[0], tri[2] - tri[0]))


def coordinate_system(n):
    """src/vector.h:532-542"""
    a = 1.0 / (1.0 + n[:, 2])
    b = -n[:, 0] * n[:, 1] * a
    x = np.stack([1.0 - n[:, 0] ** 2 * a, b, -n[:, 0]], axis=1)
    y = np.stack([b, 1.0 - n[:, 1] ** 2 * a, -n[:, 1]], axis=1)
    flip = n[:, 2] < -1.0 + 1e-6
    x[flip] = [0.0, -1.0, 0.0]
    y[flip] = [-1.0, 0.0, 0.0]
    return x, y


def lambert_bsdf(kd, n_geom, n_shade, wi, wo, two_sided):
    """src/material.h:353-449 for a material without specular lobe: diffuse * |cos_o| / pi with the reference's
    side / grazing-angle rejections."""
    gn = np.where((np.einsum("ij,ij->i", n_geom, n_shade) < 0)[:, None], -n_geom, n_geom)
    gwi, gwo = np.einsum("ij,ij->i", gn, wi), np.einsum("ij,ij->i", gn, wo)
    swi, swo = np.abs(np.einsum("ij,ij->i", n_shade, wi)), np.abs(np.einsum("ij,ij->i", n_shade, wo))
    ok = ~(gwi * gwo < 0)
    if not two_sided:
        ok &= ~((gwi < 0) & (gwo < 0))
    ok &= ~((swi == 0) | (swo <= 1e-3) | (np.abs(gwo) <= 1e-3))
    return np.where(ok[:, None], np.maximum(kd, 0.0) * (swo / np.pi)[:, None], 0.0)


def lambert_pdf(n_geom, n_shade, wi, wo, two_sided):
    """src/material.h:1023-1093 with specular weight 0 (diffuse pmf = 1)."""
    gn = np.where((np.einsum("ij,ij->i", n_geom, n_shade) < 0)[:, None], -n_geom, n_geom)
    gwi, gwo = np.einsum("ij,ij->i", gn, wi), np.einsum("ij,ij->i", gn, wo)
    swo = np.abs(np.einsum("ij,ij->i", n_shade, wo))
    ok = ~(gwi * gwo < 0)
    if not two_sided:
        ok &= ~((gwi < 0) & (gwo < 0))
    return np.where(ok, swo / np.pi, 0.0)


def render_forward(scene, spp, seed, sampler="sobol"):
    """Forward image of a diffuse scene with max_bounces = 1, following src/pathtracer.cpp:240-390 stage by stage.
    scene = dict(camera=..., shapes=[dict(vertices, indices, material_id, light_id)], materials=[dict(kd, two_sided)],
                 lights=[dict(shape_id, intensity, two_sided)])"""
    cam = scene["camera"]
    h, w = cam["resolution"]
    n = h * w
    tris, tri_shape = [], []
    for si, s in enumerate(scene["shapes"]):
        for t in s["indices"]:
            tris.append(s["vertices"][t].astype(np.float64))
            tri_shape.append(si)
    tris = np.array(tris)
    tri_shape = np.array(tri_shape)
    normals = normalize(np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]))
    # light tables, src/scene.cpp:197-253 and :38-61
    lights = scene["lights"]
    pmf, areas, cdfs, light_tris = [], [], [], []
    for li, l in enumerate(lights):
        ids = np.nonzero(tri_shape == l["shape_id"])[0]
        a = np.array([0.5 * np.linalg.norm(np.cross(tris[i][1] - tris[i][0], tris[i][2] - tris[i][0])) for i in ids])
        tot = a.sum()
        cdfs.append(np.concatenate([[0.0], np.cumsum(a)[:-1]]) / tot)
        areas.append(tot)
        lum = 0.212671 * l["intensity"][0] + 0.715160 * l["intensity"][1] + 0.072169 * l["intensity"][2]
        pmf.append(tot * lum * np.pi)
        light_tris.append(ids)
    pmf = np.array(pmf) / np.sum(pmf)
    light_cdf = np.concatenate([[0.0], np.cumsum(pmf)[:-1]])
    stream = SobolStream(seed, n) if sampler == "sobol" else PCGStream(seed, n)
    img = np.zeros((n, 3))
    weight = 1.0 / spp
    for s in range(spp):
        stream.begin_sample(s)
        org, d = primary_rays(cam, stream.next(2))
        hit, t = intersect_all(tris, org, d, 1e-3, np.inf)
        ls, bs = stream.next(4), stream.next(3)
        act = np.nonzero(hit >= 0)[0]
        if act.size == 0:
            continue
        ti = hit[act]
        # re-derived hit point, src/shape.h:289-296
        v0, v1, v2 = tris[ti, 0], tris[ti, 1], tris[ti, 2]
        e1, e2 = v1 - v0, v2 - v0
        o, dd = org[act], d[act]
        pvec = np.cross(dd, e2)
        div = np.einsum("ij,ij->i", pvec, e1)
        q = np.cross(o - v0, e1)
        tt = np.einsum("ij,ij->i", e2, q) / div
        p = o + dd * tt[:, None]
        ng = normals[ti]
        wi = -dd
        shape_of = tri_shape[ti]
        mat = [scene["materials"][scene["shapes"][si]["material_id"]] for si in shape_of]
        kd = np.array([m["kd"] for m in mat])
        two_sided = bool(mat[0].get("two_sided", False))
        # emission at the first hit, src/primary_contribution.cpp:13-26
        light_of_shape = {l["shape_id"]: li for li, l in enumerate(lights)}
        emis = np.zeros((act.size, 3))
        for k, si in enumerate(shape_of):
            if si in light_of_shape:
                l = lights[light_of_shape[si]]
                if l.get("two_sided", False) or np.dot(wi[k], ng[k]) > 0:
                    emis[k] = l["intensity"]
        img[act] += weight * emis
        # next-event estimation, src/scene.cpp:692-741 + src/path_contribution.cpp:27-50
        lsel = np.clip(np.searchsorted(light_cdf, ls[act, 0], side="right") - 1, 0, len(lights) - 1)
        nee = np.zeros((act.size, 3))
        bsdf_dir_contrib = np.zeros((act.size, 3))
        for li, l in enumerate(lights):
            m = np.nonzero(lsel == li)[0]
            if m.size == 0:
                continue
            tsel = np.clip(np.searchsorted(cdfs[li], ls[act[m], 1], side="right") - 1, 0, len(cdfs[li]) - 1)
            lt = tris[light_tris[li][tsel]]
            a = np.sqrt(ls[act[m], 2])
            b1, b2 = 1.0 - a, a * ls[act[m], 3]
            lp = lt[:, 0] + (lt[:, 1] - lt[:, 0]) * b1[:, None] + (lt[:, 2] - lt[:, 0]) * b2[:, None]
            ln = normalize(np.cross(lt[:, 1] - lt[:, 0], lt[:, 2] - lt[:, 0]))
            dirv = lp - p[m]
            dist_sq = np.einsum("ij,ij->i", dirv, dirv)
            wo = dirv / np.sqrt(dist_sq)[:, None]
            vis = ~occluded(tris, p[m], wo, 1e-3, (1 - 1e-3) * np.sqrt(dist_sq))
            facing = np.einsum("ij,ij->i", -wo, ln) > 0 if not l.get("two_sided", False) else np.ones(m.size, bool)
            f = lambert_bsdf(kd[m], ng[m], ng[m], wi[m], wo, two_sided)
            G = np.abs(np.einsum("ij,ij->i", wo, ln)) / dist_sq
            pdf_nee = pmf[li] / areas[li]
            pdf_b = lambert_pdf(ng[m], ng[m], wi[m], wo, two_sided) * G
            mis = 1.0 / (1.0 + (pdf_b / pdf_nee) ** 2)
            c = (mis * G / pdf_nee)[:, None] * f * np.asarray(l["intensity"], dtype=np.float64)
            nee[m] = np.where((vis & facing & (dist_sq > 1e-20))[:, None], c, 0.0)
        # BSDF sampling (cosine hemisphere), src/material.h:694-767, and its MIS-weighted light hit,
        # src/path_contribution.cpp:70-98
        phi = 2.0 * np.float64(np.float32(np.pi)) * bs[act, 0]
        tmp = np.sqrt(np.maximum(1.0 - bs[act, 1], 0.0))
        local = np.stack([np.cos(phi) * tmp, np.sin(phi) * tmp, np.sqrt(bs[act, 1])], axis=1)
        # shading frame of a mesh without uvs / normals: dpdu = v1 - v0 (src/shape.h:276-312 with the default uvs),
        # x = normalize(dpdu), y = normalize(n x x), x = y x n (src/shape.h:346-354)
        fx = normalize(v1 - v0)
        fy = normalize(np.cross(ng, fx))
        fx = np.cross(fy, ng)
        wdir = fx * local[:, :1] + fy * local[:, 1:2] + ng * local[:, 2:3]
        gwi = np.einsum("ij,ij->i", ng, wi)
        flip = np.einsum("ij,ij->i", ng, wdir) * gwi < 0
        wdir = np.where(flip[:, None], -wdir, wdir)
        if not two_sided:
            wdir = np.where((gwi < 0)[:, None], 0.0, wdir)
        valid_dir = np.einsum("ij,ij->i", wdir, wdir) > 1e-3
        bh, _ = intersect_all(tris, p, wdir, 1e-3, np.inf)
        for k in np.nonzero((bh >= 0) & valid_dir)[0]:
            si = tri_shape[bh[k]]
            bt = tris[bh[k]]
            e1b, e2b = bt[1] - bt[0], bt[2] - bt[0]
            pv = np.cross(wdir[k], e2b)
            dv = pv @ e1b
            qv = np.cross(p[k] - bt[0], e1b)
            tb = (e2b @ qv) / dv
            bp = p[k] + wdir[k] * tb
            dirv = bp - p[k]
            dist_sq = dirv @ dirv
            wo = dirv / np.sqrt(dist_sq)
            pdf_b = lambert_pdf(ng[k:k + 1], ng[k:k + 1], wi[k:k + 1], wo[None], two_sided)[0]
            if not (dist_sq > 1e-20 and pdf_b > 1e-20) or si not in light_of_shape:
                continue
            li = light_of_shape[si]
            l = lights[li]
            nb = normals[bh[k]]
            if not (l.get("two_sided", False) or np.dot(-wo, nb) > 0):
                continue
            f = lambert_bsdf(kd[k:k + 1], ng[k:k + 1], ng[k:k + 1], wi[k:k + 1], wo[None], two_sided)[0]
            G = abs(np.dot(wo, nb)) / dist_sq
            pdf_nee = (pmf[li] / areas[li]) / G
            mis = 1.0 / (1.0 + (pdf_nee / pdf_b) ** 2)
            bsdf_dir_contrib[k] = (mis / pdf_b) * f * np.asarray(l["intensity"], dtype=np.float64)
        img[act] += weight * (nee + bsdf_dir_contrib)
    return img.reshape(h, w, 3)
