"""Synthetic workloads for the modal-synthesis path (SURVEY.md section 8(d) conventions).

The reference ships no .modes / material / .fatcube data, so every test
and benchmark input is generated here, seeded, with numpy only.  Nothing in this module computes the
synthesis itself.
"""
import numpy as np

SAMPLE_RATE = 44100              # reference config.h:13
H = 1.0 / SAMPLE_RATE
BUF = 256                        # BASELINE.json buffer size (ModalSolver<double,256>)
SPEED_OF_SOUND = 343.0

MATERIALS = {
    # density, youngsModulus, poissonRatio, alpha, beta   (ModalMaterial.h:20-26)
    "low_damping": dict(density=2600.0, youngsModulus=6.2e10, poissonRatio=0.2, alpha=1.0, beta=1e-7),
    "high_damping": dict(density=2600.0, youngsModulus=6.2e10, poissonRatio=0.2, alpha=30.0, beta=5e-7),
}


def mode_frequencies(M, seed, fmin=80.0, fmax=18000.0):
    """Log-uniform mode frequencies in [fmin, fmax] Hz, ascending."""
    rng = np.random.default_rng(seed)
    return np.sort(np.exp(rng.uniform(np.log(fmin), np.log(fmax), M)))


def omega_squared(freqs, density):
    """Inverse of modal_integrator.h:63: omega = sqrt(omega2/density) = 2 pi f."""
    return density * (2.0 * np.pi * np.asarray(freqs)) ** 2


def mode_shapes(M, K, seed, dtype=np.float64):
    """U[m][d] ~ N(0,1), mode-major like ModeData::_modes (ModeData.h:23-24)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((M, K)).astype(dtype)


def ffat_geometry(R=1.5, n=32, center=(0.0, 0.0, 0.0)):
    """One cube-map geometry (shell #2 fields kept by ffat_map_serialize.h:71-78).  Face order
    +x,-x,+y,-y,+z,-z; lowCorners follow the quad-corner convention of ffat_solver.h:365-401,420."""
    c = np.asarray(center, dtype=np.float64)
    low = np.empty((6, 3))
    for f in range(6):
        dk = f // 2
        corner = c - R
        if f % 2 == 0:
            corner = corner.copy(); corner[dk] = c[dk] + R
        low[f] = corner
    return dict(cellsize=2.0 * R / n, lowcorners=low,
                n_elements=np.full((6, 2), n, dtype=np.int32),
                strides=(np.arange(6) * n * n).astype(np.int32),
                center1=c.copy(), bboxlow=c - R, bboxtop=c + R, center=c.copy())


def texel_centres(geom):
    """World positions of all texel centres, in Psi index order (GetDataQuadStride, ffat_solver.h:141-144)."""
    h = geom["cellsize"]; pts = []
    for f in range(6):
        dk = f // 2; di = (dk + 1) % 3; dj = (dk + 2) % 3
        Nx, Ny = geom["n_elements"][f]
        x, y = np.meshgrid(np.arange(Nx), np.arange(Ny), indexing="ij")
        p = np.empty((Nx * Ny, 3))
        p[:, dk] = geom["lowcorners"][f][dk]
        p[:, di] = geom["lowcorners"][f][di] + (x.ravel() + 0.5) * h
        p[:, dj] = geom["lowcorners"][f][dj] + (y.ravel() + 0.5) * h
        pts.append(p)
    return np.concatenate(pts)


def ffat_maps(freqs, seed0=2000, R=1.5, n=32, geom=None):
    """One map dict per mode: shared geometry, k_m = 2 pi f_m / 343, Psi_m a positive smooth field
    (|low-order harmonic mix| + 0.1) seeded seed0 + m."""
    geom = ffat_geometry(R, n) if geom is None else geom
    pts = texel_centres(geom) - geom["center"]
    u = pts / np.linalg.norm(pts, axis=1, keepdims=True)
    x, y, z = u[:, 0], u[:, 1], u[:, 2]
    basis = np.stack([np.ones_like(x), x, y, z, x * y, y * z, z * x, x * x - y * y, 3 * z * z - 1])
    maps = []
    for m, f in enumerate(freqs):
        c = np.random.default_rng(seed0 + m).standard_normal(basis.shape[0])
        psi = np.abs(c @ basis) + 0.1
        d = dict(geom); d.update(k=2.0 * np.pi * float(f) / SPEED_OF_SOUND, psi=psi, modeid=m,
                                 is_compressed=False)
        maps.append(d)
    return maps


def cubemap_vertices(center, half_cells, cell_size):
    """Quad vertices of one cube-map shell in the order FFAT_Map<T,1>::CubemapMesh emits them (reference
    ffat_solver.h:334-397): faces +x,-x,+y,-y,+z,-z; per face ii over axis di=(dk+1)%3, jj over dj=(dk+2)%3;
    four vertices per quad, offsets (-,-),(+,-),(+,+),(-,+) * cell/2 about the cell centre.  The box is
    2*half_cells[axis] cells wide (half_cells: int or per-axis triple) and centred on `center`.
    Returns (V [4*n_quads][3], n_elements [6][2])."""
    center = np.asarray(center, dtype=np.float64); h = float(cell_size)
    hc = np.broadcast_to(np.asarray(half_cells, dtype=np.int64), (3,))
    n = 2 * hc
    low = center - hc * h
    off = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=np.float64) * h / 2.0
    V = []; ne = []
    for face in range(6):
        dk = face // 2; di = (dk + 1) % 3; dj = (dk + 2) % 3
        plane = low[dk] + (n[dk] * h if face % 2 == 0 else 0.0)
        ii, jj = np.meshgrid(np.arange(n[di]), np.arange(n[dj]), indexing="ij")
        cc = np.empty((n[di], n[dj], 4, 3))
        cc[..., dk] = plane
        cc[..., di] = (low[di] + (0.5 + ii) * h)[..., None] + off[:, 0]
        cc[..., dj] = (low[dj] + (0.5 + jj) * h)[..., None] + off[:, 1]
        V.append(cc.reshape(-1, 3)); ne.append([n[di], n[dj]])
    return np.concatenate(V), np.asarray(ne, dtype=np.int32)


def ffat_fit_workload(n_maps, seed, half_cells=(8, 12, 16), cell_size=0.09375, center=(0.0, 0.0, 0.0), noise=0.0):
    """Synthetic input of FFAT_Map<T,3>::Solve: three nested cube shells (shell 2 outermost, the one the
    run-time map keeps), one wavenumber per mode and the Dirichlet pressure of a far-field radiator
    p(x) = Psi(dir) e^{-ikr} / (k r) sampled at the quad centres (two entries per quad like the reference's
    triangle-mesh ordering; the odd ones hold a different value on purpose: Solve must skip them).
    Returns dict(V, n_elements [S][6][2], cell_size, k [n_maps], pressure complex [n_maps][2*n_total], psi_true)."""
    rng = np.random.default_rng(seed)
    Vs, nes, cents = [], [], []
    for hc in half_cells:
        V, ne = cubemap_vertices(center, hc, cell_size)
        Vs.append(V); nes.append(ne)
        cents.append(V.reshape(-1, 4, 3).mean(axis=1))
    X = np.concatenate(cents) - np.asarray(center)
    r = np.linalg.norm(X, axis=1); u = X / r[:, None]
    x, y, z = u[:, 0], u[:, 1], u[:, 2]
    basis = np.stack([np.ones_like(x), x, y, z, x * y, y * z, z * x, x * x - y * y, 3 * z * z - 1])
    freqs = mode_frequencies(n_maps, seed)
    k = 2.0 * np.pi * freqs / SPEED_OF_SOUND
    n_total = len(r)
    P = np.empty((n_maps, 2 * n_total), dtype=np.complex128)
    psi_true = np.empty((n_maps, n_total))
    for m in range(n_maps):
        c = rng.standard_normal(basis.shape[0])
        psi = np.abs(c @ basis) + 0.1
        p = psi * np.exp(-1j * k[m] * r) / (k[m] * r)
        if noise:
            p = p * (1.0 + noise * rng.standard_normal(n_total))
        P[m, 0::2] = p
        P[m, 1::2] = -3.0 * p + 1.0
        psi_true[m] = psi
    return dict(V=np.concatenate(Vs), n_elements=np.stack(nes), cell_size=float(cell_size), k=k, pressure=P,
                psi_true=psi_true, n_total=n_total)


def listeners(L, seed, rmin=3.0, rmax=10.0):
    """Listener positions outside the bbox, directions uniform on S^2 (never axis-aligned)."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((L, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d * rng.uniform(rmin, rmax, (L, 1))


def unit_vectors(n, seed):
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3))
    return d / np.linalg.norm(d, axis=1, keepdims=True)


def ab_from_material(freqs, mat):
    """(a, b) of  q'' + a q' + b q = f  for given mode frequencies -- what ModalIntegrator::Build
    (modal_integrator.h:47-70) derives from omega^2 = density (2 pi f)^2."""
    omega = 2.0 * np.pi * np.asarray(freqs, dtype=np.float64)
    xi = 0.5 * (mat["alpha"] / omega + mat["beta"] * omega)
    return 2.0 * xi * omega, omega ** 2


def batch_workload(n_obj, n_modes, n_buf, seed, material="low_damping", first_second_bufs=None):
    """cfg5-style offline batch: per-object mode sets, one PointForce per object at a seeded random
    buffer within the first second, one static listener.  Returns dict of float64 arrays:
    a, b [n_obj][n_modes]; space [n_obj][n_modes] (modal load U^T f); trans [n_obj][n_modes];
    imp_buf [n_obj]."""
    rng = np.random.default_rng(seed)
    mat = MATERIALS[material]
    f = np.sort(np.exp(rng.uniform(np.log(80.0), np.log(18000.0), (n_obj, n_modes))), axis=1)
    a, b = ab_from_material(f, mat)
    space = rng.standard_normal((n_obj, n_modes))
    k = 2.0 * np.pi * f / SPEED_OF_SOUND
    r = rng.uniform(3.0, 10.0, (n_obj, 1))
    trans = (np.abs(rng.standard_normal((n_obj, n_modes))) + 0.1) / (k * r)
    if first_second_bufs is None:
        first_second_bufs = max(1, min(n_buf, SAMPLE_RATE // BUF))
    imp = rng.integers(0, first_second_bufs, n_obj).astype(np.int32)
    return dict(a=a, b=b, space=space, trans=trans, imp_buf=imp, freqs=f)
