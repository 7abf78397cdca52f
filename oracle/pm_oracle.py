"""CPU restatement (numpy, float64) of MG-PICOLA's per-step COLA particle-mesh force path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product library imports this module; it may be used by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.

Every function cites the reference lines (relative to the reference's src/) it restates.  The
restatement is pinned against the unmodified reference compiled into oracle/_ref (see
oracle/Makefile, oracle/ref_lib.py) by tests/test_oracle_vs_ref.py and against the committed
fixtures in tests/golden/ (generated from oracle/_ref by oracle/make_golden.py).

Particle data are float32 and grids float64, i.e. the reference's default build
(MEMORY_MODE on, SINGLE_PRECISION off; Makefile:121-122, 155-156); pass grid_dtype=np.float32 for
the -DSINGLE_PRECISION variant.  Single task (NTask = 1): Local_nx = Nmesh, Local_x_start = 0.
"""
import numpy as np
import scipy.fft as sfft

INVERSE_H0_MPCH = 2997.92458   # vars.h:64
PI = 3.14159265358979323846    # vars.h:60


# ----------------------------------------------------------------------------- CIC

def _cic(pos, nmesh, box, wpar=1.0):
    """Cell indices and weights exactly as auxPM.c:298-330 / 576-603 compute them."""
    scale = np.float64(nmesh) / np.float64(box)
    X = pos[:, 0].astype(np.float64) * scale
    Y = pos[:, 1].astype(np.float64) * scale
    Z = pos[:, 2].astype(np.float64) * scale
    IX = X.astype(np.uint32).astype(np.int64)
    IY = Y.astype(np.uint32).astype(np.int64)
    IZ = Z.astype(np.uint32).astype(np.int64)
    DX = X - IX
    DY = Y - IY
    DZ = Z - IZ
    TX = 1.0 - DX
    TY = 1.0 - DY
    TZ = 1.0 - DZ
    DY = DY * wpar
    TY = TY * wpar
    IY[IY >= nmesh] = 0
    IZ[IZ >= nmesh] = 0
    IXn = IX + 1                      # no wrap: ghost slice on the right (auxPM.c:325-326)
    IYn = IY + 1
    IZn = IZ + 1
    IYn[IYn >= nmesh] = 0
    IZn[IZn >= nmesh] = 0
    return (IX, IY, IZ, IXn, IYn, IZn, DX, DY, DZ, TX, TY, TZ)


def ptomesh_deposit(pos, nmesh, nsample, box, grid_dtype=np.float64):
    """PtoMesh up to (not including) the FFT: auxPM.c:288-356.

    Returns the padded real grid [Nmesh][Nmesh][2*(Nmesh/2+1)] holding delta (the ghost slice has
    been folded into slice 0 as the self-Sendrecv of a single task does)."""
    N = nmesh
    nzp = 2 * (N // 2 + 1)
    wpar = (np.float64(N) / np.float64(nsample)) ** 3          # auxPM.c:289
    IX, IY, IZ, IXn, IYn, IZn, DX, DY, DZ, TX, TY, TZ = _cic(pos, N, box, wpar)
    size = (N + 1) * N * nzp
    grid = np.zeros(size, dtype=np.float64)

    def add(ix, iy, iz, w):
        idx = (ix * N + iy) * nzp + iz
        grid[:] += np.bincount(idx, weights=w, minlength=size)

    add(IX, IY, IZ, TX * TY * TZ)        # auxPM.c:335-342
    add(IX, IY, IZn, TX * TY * DZ)
    add(IX, IYn, IZ, TX * DY * TZ)
    add(IX, IYn, IZn, TX * DY * DZ)
    add(IXn, IY, IZ, DX * TY * TZ)
    add(IXn, IY, IZn, DX * TY * DZ)
    add(IXn, IYn, IZ, DX * DY * TZ)
    add(IXn, IYn, IZn, DX * DY * DZ)
    grid = (grid - 1.0).astype(grid_dtype).reshape(N + 1, N, nzp)   # density starts at -1 (auxPM.c:292)
    # ghost slice -> slice 0:  density[i] += temp[i] + 1.0   (auxPM.c:350-355)
    grid[0] += (grid[N] + grid_dtype(1.0)).astype(grid_dtype)
    return np.ascontiguousarray(grid[:N])


def r2c(grid_real, nmesh):
    """my_fftw_execute on an r2c plan (wrappers.c:42-46): unnormalised, half spectrum [kx][ky][kz<=N/2]."""
    N = nmesh
    out = sfft.rfftn(grid_real[:, :, :N].astype(np.float64), workers=-1)
    return out


def c2r(grid_k, nmesh):
    """Unnormalised c2r over a half spectrum that need not be Hermitian on the kz = 0, N/2 planes:
    c2c over x, y first, c2r over z last (FFTW and cuFFT order); irfftn does exactly this."""
    N = nmesh
    return sfft.irfftn(grid_k, s=(N, N, N), axes=(0, 1, 2), workers=-1) * (np.float64(N) ** 3)


def _dvec(nmesh):
    N = nmesh
    i = np.arange(N)
    d = np.where(i > N // 2, i - N, i).astype(np.float64)       # auxPM.c:478: iglobal > Nmesh/2 ? iglobal-Nmesh
    dz = np.arange(N // 2 + 1).astype(np.float64)
    return d[:, None, None], d[None, :, None], dz[None, None, :]


def forces_kspace(P3D, nmesh, box, mg_phik=None):
    """Forces() before the inverse FFTs: auxPM.c:450-535.  Returns FN11, FN12, FN13 (complex)."""
    N = nmesh
    if mg_phik is not None:
        P3D = P3D + mg_phik                                      # auxPM.c:450-455
    d0, d1, d2 = _dvec(N)
    RK = d0 * d0 + d1 * d1 + d2 * d2
    with np.errstate(divide="ignore"):
        KK = -1.0 / RK
    scale = 2.0 * np.pi / box                                    # Scale (auxPM.c:441)
    n3 = np.float64(N) ** 3
    out = []
    with np.errstate(invalid="ignore"):
        dens0 = (P3D.real * KK) / n3                             # auxPM.c:497-498
        dens1 = (-1.0 * P3D.imag * KK) / n3
        for d in (d0, d1, d2):
            f = (dens1 * d / scale) + 1j * (dens0 * d / scale)   # auxPM.c:503-508
            f[0, 0, 0] = 0.0                                     # auxPM.c:465-470
            out.append(f)
    return out


def forces(P3D, nmesh, box, mg_phik=None):
    """Forces(): k-space kernel + 3 c2r (auxPM.c:437-555).  Returns N11, N12, N13 real [N][N][N]."""
    return [c2r(f, nmesh) for f in forces_kspace(P3D, nmesh, box, mg_phik)]


def mtoparticles(pos, N11, N12, N13, nmesh, box, tot_numpart=None):
    """MtoParticles: auxPM.c:560-644.  Disp is float32 (MEMORY_MODE), sums in double."""
    N = nmesh
    IX, IY, IZ, IXn, IYn, IZn, DX, DY, DZ, TX, TY, TZ = _cic(pos, N, box, 1.0)
    IXn = IXn % N      # ghost slice == slice 0 of the (only) task (auxPM.c:546-551)
    disp = np.empty((pos.shape[0], 3), dtype=np.float32)
    for a, F in enumerate((N11, N12, N13)):
        v = (F[IX, IY, IZ] * TX * TY * TZ + F[IX, IY, IZn] * TX * TY * DZ +
             F[IX, IYn, IZ] * TX * DY * TZ + F[IX, IYn, IZn] * TX * DY * DZ +
             F[IXn, IY, IZ] * DX * TY * TZ + F[IXn, IY, IZn] * DX * TY * DZ +
             F[IXn, IYn, IZ] * DX * DY * TZ + F[IXn, IYn, IZn] * DX * DY * DZ)
        disp[:, a] = v.astype(np.float32)
    tot = pos.shape[0] if tot_numpart is None else tot_numpart
    sumD = disp.astype(np.float64).sum(axis=0) / np.float64(tot)   # auxPM.c:632-640
    return disp, sumD


# ----------------------------------------------------------------------------- Kick / Drift

def kick(vel, disp, D, D2, sumDxyz, omega, use_cola, A, dda, ddDddy, ddD2ddy, tot_numpart=None):
    """Kick particle loop, non-SCALEDEPENDENT branch: main.c:721-739.  Mutates nothing; returns
    (vel_new float32, disp_new float32 [mean-subtracted, as the reference leaves it], sumxyz)."""
    disp_new = (disp.astype(np.float64) - np.asarray(sumDxyz, dtype=np.float64)[None, :]).astype(np.float32)
    force = (-1.5 * omega) * disp_new.astype(np.float64) - \
        (np.float64(use_cola) * (D.astype(np.float64) * ddDddy + D2.astype(np.float64) * ddD2ddy)) / A
    vel_new = (vel.astype(np.float64) + force * dda).astype(np.float32)
    tot = vel.shape[0] if tot_numpart is None else tot_numpart
    sumxyz = vel_new.astype(np.float64).sum(axis=0) / np.float64(tot)
    return vel_new, disp_new, sumxyz


def periodic_wrap(x, box):
    """auxPM.c:649-655 in float arithmetic (x: float32 array)."""
    x = x.astype(np.float32).copy()
    b = np.float32(box)
    for _ in range(1000):
        m = x >= b
        if not m.any():
            break
        x[m] = x[m] - b
    for _ in range(1000):
        m = x < 0
        if not m.any():
            break
        x[m] = x[m] + b
    x[x == b] = np.float32(0.0)
    return x


def drift(pos, vel, D, D2, sumxyz, box, use_cola, dyyy, deltaD, deltaD2):
    """Drift particle loop, non-SCALEDEPENDENT branch: main.c:773-783."""
    p = (pos.astype(np.float64) + (vel.astype(np.float64) - np.asarray(sumxyz, dtype=np.float64)[None, :]) * dyyy
         ).astype(np.float32)
    arg = p.astype(np.float64) + np.float64(use_cola) * (D.astype(np.float64) * deltaD + D2.astype(np.float64) * deltaD2)
    return periodic_wrap(arg.astype(np.float32), box)


# ----------------------------------------------------------------------------- modified gravity (mg.h)

def _rk(nmesh):
    d0, d1, d2 = _dvec(nmesh)
    return d0 * d0 + d1 * d1 + d2 * d2


def divide_by_laplacian(P3D, nmesh, box, omega, a):
    """DivideByLaplacian: mg.h:21-64."""
    N = nmesh
    normfactor = 1.0 / np.float64(N) ** 3
    normfactor *= 1.5 * omega / a * (box / INVERSE_H0_MPCH / (2.0 * PI)) ** 2
    RK = _rk(N)
    with np.errstate(divide="ignore"):
        KK = -1.0 / RK
    with np.errstate(invalid="ignore"):
        out = normfactor * P3D * KK
    out[0, 0, 0] = 0.0
    return out


def eff_density_to_phik(dk, nmesh, coupling, massterm2):
    """EffDensitykToPhiofk: mg.h:74-116."""
    RK = _rk(nmesh)
    KK = RK / (RK + massterm2)
    out = coupling * dk * KK
    out[0, 0, 0] = 0.0
    return out


def fofr_scalars(a, omega, box, fofr0, nfofr):
    """phicrit (udf:731), coupling = 2 beta^2 with beta = 1/sqrt(6) (udf:462-470, 587-591),
    massterm2 (mg.h:80 with mass2_of_a udf:490-498)."""
    phicrit = 1.5 * fofr0 * ((omega + 4.0 * (1.0 - omega)) / (1.0 / (a * a * a) * omega + 4.0 * (1.0 - omega))) ** (nfofr + 1.0)
    beta = 1.0 / np.sqrt(6.0)
    coupling = 2.0 * beta * beta
    a3 = a * a * a
    fac = omega / a3 + 4.0 * (1.0 - omega)
    fac0 = omega + 4.0 * (1.0 - omega)
    mass2 = fac0 * (fac / fac0) ** (nfofr + 2.0) / ((1.0 + nfofr) * fofr0)
    massterm2 = a ** 2 * mass2 / ((2.0 * PI) * INVERSE_H0_MPCH / box) ** 2
    return phicrit, coupling, massterm2


def fifth_force_potential_screening(P3D, dens_real, nmesh, box, omega, a, phicrit, coupling, massterm2, screening=True):
    """ComputeFifthForce_PotentialScreening: mg.h:147-189.  dens_real = copy of delta(x) [N][N][N].
    Returns phi_k to be added to P3D in Forces (= P3D_mgarray_two)."""
    if not screening:
        return eff_density_to_phik(P3D, nmesh, coupling, massterm2)
    phik = divide_by_laplacian(P3D, nmesh, box, omega, a)
    phi = c2r(phik, nmesh)
    s = np.ones_like(phi)
    neg = ~(phi >= 0.0)
    sf = np.abs(phicrit / phi[neg])
    s[neg] = np.minimum(sf, 1.0)                                  # udf:725-737
    deff = dens_real * s                                          # mg.h:174-176
    deffk = sfft.rfftn(deff, workers=-1)
    return eff_density_to_phik(deffk, nmesh, coupling, massterm2)


def dgp_scalars(a, omega, rcH0):
    """beta_DGP (udf:453-455 with LCDM hubble udf:427, 445), coupling (udf:593-595), fac0 (udf:762)."""
    H = np.sqrt(omega / (a * a * a) + 1.0 - omega)
    dH = 1.0 / (2.0 * H) * (-3.0 * omega / (a * a * a * a))
    beta = 1.0 + 2.0 * rcH0 * (H + a * dH / 3.0)
    coupling = 1.0 / (3.0 * beta)
    fac0 = 8.0 / 9.0 * omega * (rcH0 / beta) ** 2
    return coupling, fac0


def fifth_force_density_screening(P3D, dens_real, nmesh, box, coupling, fac0, rsmooth, screening=True):
    """ComputeFifthForce_DensityScreening with the Gaussian filter: mg.h:197-260, 268-309, 318-322."""
    N = nmesh
    if not screening:
        return P3D * coupling                                      # mg.h:212-217 (density aliases P3D)
    RK = _rk(N)
    kR = np.sqrt(RK) * 2.0 * PI / box * rsmooth
    smooth = np.exp(-0.5 * kR * kR) * (1.0 / np.float64(N) ** 3)
    ds = c2r(P3D * smooth, N)
    fac = fac0 * (1.0 + ds)                                       # udf:762-765
    with np.errstate(invalid="ignore", divide="ignore"):
        sf = np.where(fac < 1e-5, 1.0, 2.0 * (np.sqrt(1.0 + fac) - 1.0) / fac)
    deff = dens_real * (coupling * sf)                            # mg.h:243
    return sfft.rfftn(deff, workers=-1)


# ----------------------------------------------------------------------------- P(k) (compute_pofk.c)

def adjust_pofk_parameters(nmesh, box, nbins, bintype, subtract_shotnoise, kmin_hmpc, kmax_hmpc):
    """compute_pofk.c:86-104 + 758-805; k limits returned in integer-k units."""
    kmin = kmin_hmpc * box / (2.0 * np.pi)
    kmax = kmax_hmpc * box / (2.0 * np.pi)
    if bintype not in (0, 1):
        bintype = 0
    if subtract_shotnoise not in (0, 1):
        subtract_shotnoise = 1
    if nbins <= 0:
        nbins = nmesh
    if kmax <= kmin:
        kmin = 0.0 if bintype == 0 else 1.0
        kmax = float(nmesh)
    if kmin < 0.0:
        kmin = 0.0 if bintype == 0 else 1.0
    if bintype == 1 and kmin == 0.0:
        kmin = 1.0
    if kmax > np.sqrt(3.0) * nmesh:
        kmax = float(nmesh)
    return nbins, bintype, subtract_shotnoise, kmin, kmax


def pofk_bin_index(kmag, kmin, kmax, nbins, bintype):
    """compute_pofk.c:30-46 (C truncation of the (int) cast)."""
    if bintype == 0:
        return np.trunc((kmag - kmin) / (kmax - kmin) * nbins + 0.5).astype(np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        idx = np.trunc(np.log(kmag / kmin) / np.log(kmax / kmin) * nbins + 0.5)
    idx = np.where(kmag <= 0.0, -1, idx)
    return idx.astype(np.int64)


def compute_power_spectrum(P3D, nmesh, nsample, box, nbins, bintype, subtract_shotnoise, kmin_hmpc, kmax_hmpc):
    """compute_power_spectrum: compute_pofk.c:71-236.  Returns (pofk, k_mean, n_modes) per bin."""
    N = nmesh
    nbins, bintype, shot, kmin, kmax = adjust_pofk_parameters(N, box, nbins, bintype, subtract_shotnoise, kmin_hmpc, kmax_hmpc)
    i = np.arange(N)
    dxy = np.where(i > N // 2, N - i, i)                           # |kx|, |ky| (compute_pofk.c:127, 135 + mirror rows)
    dz = np.arange(N // 2 + 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        gxy = np.where(dxy == 0, 1.0, np.sin((PI * dxy) / float(N)) / ((PI * dxy) / float(N)))
        gz = np.where(dz == 0, 1.0, np.sin((PI * dz) / float(N)) / ((PI * dz) / float(N)))
    gz[N // 2] = 2.0 / PI                                          # compute_pofk.c:169
    kmag = np.sqrt((dxy[:, None, None] ** 2 + dxy[None, :, None] ** 2 + dz[None, None, :] ** 2).astype(np.float64))
    nk = pofk_bin_index(kmag, kmin, kmax, nbins, bintype)
    corr = 1.0 / (gxy[:, None, None] * gxy[None, :, None] * gz[None, None, :]) ** 4.0 * (1.0 / np.float64(N) ** 6)
    p = (P3D.real * P3D.real + P3D.imag * P3D.imag) * corr
    w = np.full(N // 2 + 1, 2.0)
    w[0] = 1.0
    w[N // 2] = 1.0
    w = np.broadcast_to(w[None, None, :], p.shape)
    ok = (nk >= 0) & (nk < nbins)
    pofk_bin = np.bincount(nk[ok], weights=(w * p)[ok], minlength=nbins)
    k_bin = np.bincount(nk[ok], weights=(w * kmag)[ok], minlength=nbins)
    n_bin = np.bincount(nk[ok], weights=w[ok], minlength=nbins)
    pofk = np.zeros(nbins)
    kmean = np.zeros(nbins)
    good = n_bin > 0
    pofk[good] = pofk_bin[good] / n_bin[good] * box ** 3
    if shot:
        pofk[good] -= (box / float(nsample)) ** 3
    kmean[good] = k_bin[good] / n_bin[good] * 2.0 * np.pi / box
    return pofk, kmean, n_bin


# ----------------------------------------------------------------------------- one full step

def get_displacements(pos, nmesh, nsample, box, model="none", mg=None, grid_dtype=np.float64, pofk=None):
    """GetDisplacements for a single task: auxPM.c:37-103.  mg = dict of per-step scalars.
    Returns dict(disp, sumDxyz, density_k, pofk)."""
    N = nmesh
    dens = ptomesh_deposit(pos, N, nsample, box, grid_dtype)
    dens_real = dens[:, :, :N].astype(np.float64)
    P3D = r2c(dens, N)
    out = {}
    if pofk is not None:
        out["pofk"] = compute_power_spectrum(P3D, N, nsample, box, **pofk)
    phik = None
    if model == "fofr":
        phik = fifth_force_potential_screening(P3D, dens_real, N, box, mg["omega"], mg["a"], mg["phi_crit"],
                                               mg["coupling"], mg["massterm2"], mg.get("screening", True))
    elif model == "dgp":
        phik = fifth_force_density_screening(P3D, dens_real, N, box, mg["coupling"], mg["dgp_fac0"], mg["rsmooth"],
                                             mg.get("screening", True))
    elif model == "geff":
        P3D = P3D * mg["geff"]
    N11, N12, N13 = forces(P3D, N, box, phik)
    disp, sumD = mtoparticles(pos, N11, N12, N13, N, box)
    out.update(disp=disp, sumDxyz=sumD, density_k=P3D, force_grids=(N11, N12, N13), density=dens)
    return out


# ----------------------------------------------------------------------------- scale-dependent growth (2LPT.c:1539-2005)

FIELD_D, FIELD_dDdy, FIELD_ddDddy, FIELD_deltaD = 0, 1, 2, 3      # proto.h:155-158


def sd_ivec(nmesh):
    """Integer wave vector with the IC code's Nyquist convention idx < N/2 ? idx : idx - N (2LPT.c:1589-1605)."""
    N = nmesh
    i = np.arange(N)
    d = np.where(i < N // 2, i, i - N).astype(np.int64)
    k = np.arange(N // 2 + 1)
    dz = np.where(k < N // 2, k, k - N).astype(np.int64)
    return d[:, None, None], d[None, :, None], dz[None, None, :]


def sd_k_of_m(nmesh, box):
    """|k| in h/Mpc for every integer m = |d|^2 in [0, 3 (N/2)^2]: the argument at which the driver evaluates
    growth_X_scaledependent when it fills the table that crosses the C ABI."""
    m = np.arange(3 * (nmesh // 2) ** 2 + 1, dtype=np.float64)
    return 2.0 * PI / box * np.sqrt(m)


def sd_field_kspace(delta_k, growth_by_k2, lpt_order, nmesh, box):
    """from_cdisp_store_to_ZA, the k-space loop (2LPT.c:1584-1630): the three components
    cdisp_D[a] = (-Im d, Re d) * kvec_a / kmag2 * growth_factor, growth_factor = normfactor * G(|k|)."""
    N = nmesh
    d0, d1, d2 = sd_ivec(N)
    m = d0 * d0 + d1 * d1 + d2 * d2
    normfactor = 1.0 if lpt_order == 1 else -3.0 / 7.0 / np.float64(N) ** 3
    g = normfactor * np.asarray(growth_by_k2, dtype=np.float64)[m]
    kv = [d * 2 * PI / box for d in (d0, d1, d2)]
    kmag2 = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2]
    out = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for a in range(3):
            re = -delta_k.imag * kv[a] / kmag2 * g
            im = delta_k.real * kv[a] / kmag2 * g
            f = re + 1j * im
            f[0, 0, 0] = 0.0
            out.append(f)
    return out


def lagrangian_readout(grids, nmesh, nsample):
    """Trilinear read-out of real grids [N][N][N] at the Lagrangian lattice (2LPT.c:1657-1705), single task.
    Returns [Nsample^3][len(grids)] doubles in the order coord = (n Ns + m) Ns + p."""
    N, Ns = nmesh, nsample
    n = np.arange(Ns)
    u = (n * N).astype(np.float64) / np.float64(Ns)
    i = u.astype(np.int64)
    i[i == N] = N - 1
    fu = u - i
    ii = (i + 1) % N                         # x: ghost slice == slice 0 of the single task; y, z wrap (2LPT.c:1679-1680)
    I, J, K = np.meshgrid(i, i, i, indexing="ij")
    II, JJ, KK = np.meshgrid(ii, ii, ii, indexing="ij")
    U, V, W = np.meshgrid(fu, fu, fu, indexing="ij")
    f1 = (1 - U) * (1 - V) * (1 - W)
    f2 = (1 - U) * (1 - V) * W
    f3 = (1 - U) * V * (1 - W)
    f4 = (1 - U) * V * W
    f5 = U * (1 - V) * (1 - W)
    f6 = U * (1 - V) * W
    f7 = U * V * (1 - W)
    f8 = U * V * W
    out = []
    for G in grids:
        v = (G[I, J, K] * f1 + G[I, J, KK] * f2 + G[I, JJ, K] * f3 + G[I, JJ, KK] * f4 +
             G[II, J, K] * f5 + G[II, J, KK] * f6 + G[II, JJ, K] * f7 + G[II, JJ, KK] * f8)
        out.append(v.reshape(-1))
    return np.stack(out, axis=1)


def sd_displacement_field(delta_k, growth_by_k2, lpt_order, nmesh, nsample, box, grid_dtype=np.float64):
    """from_cdisp_store_to_ZA (2LPT.c:1539-1735): returns ZA_D[Nsample^3][3] (float_kind), mean removed.
    assign_displacment_field_to_particles then copies ZA_D[coord_q] into the float particle field."""
    N = nmesh
    ck = np.complex64 if grid_dtype == np.float32 else np.complex128
    comps = [c2r(f.astype(ck).astype(np.complex128), N).astype(grid_dtype).astype(np.float64)
             for f in sd_field_kspace(delta_k, growth_by_k2, lpt_order, N, box)]
    za = lagrangian_readout(comps, N, nsample)
    mean = za.sum(axis=0) / np.float64(nsample) ** 3                     # 2LPT.c:1716-1719
    za = za.astype(grid_dtype)                                            # ZA_D is float_kind (2LPT.c:1700)
    return (za.astype(np.float64) - mean[None, :]).astype(grid_dtype)     # 2LPT.c:1724-1728


def kick_sd(vel, disp, D, D2, sumDxyz, omega, use_cola, A, dda, tot_numpart=None):
    """Kick, SCALEDEPENDENT branch (main.c:705-717): P.D / P.D2 hold the ddD-weighted fields; their sum and the
    product with the int UseCOLA are float arithmetic in C."""
    disp_new = (disp.astype(np.float64) - np.asarray(sumDxyz, dtype=np.float64)[None, :]).astype(np.float32)
    dd = (np.float32(use_cola) * (D.astype(np.float32) + D2.astype(np.float32))).astype(np.float32)
    force = (-1.5 * omega) * disp_new.astype(np.float64) - dd.astype(np.float64) / A
    vel_new = (vel.astype(np.float64) + force * dda).astype(np.float32)
    tot = vel.shape[0] if tot_numpart is None else tot_numpart
    return vel_new, disp_new, vel_new.astype(np.float64).sum(axis=0) / np.float64(tot)


def drift_sd(pos, vel, dDdy, dD2dy, sumxyz, box, use_cola, dyyy):
    """Drift, SCALEDEPENDENT branch (main.c:760-770): P.dDdy / P.dD2dy hold the per-particle increments; the
    second statement is float arithmetic throughout."""
    p = (pos.astype(np.float64) + (vel.astype(np.float64) - np.asarray(sumxyz, dtype=np.float64)[None, :]) * dyyy
         ).astype(np.float32)
    arg = (p + (np.float32(use_cola) * (dDdy.astype(np.float32) + dD2dy.astype(np.float32))).astype(np.float32)).astype(np.float32)
    return periodic_wrap(arg, box)


def init_particles_sd(D, D2, dDdy, dD2dy, nsample, box, use_cola):
    """main.c:257-304, SCALEDEPENDENT branch: returns (pos, vel, ids) in Lagrangian order."""
    Ns = nsample
    q = np.stack(np.meshgrid(np.arange(Ns), np.arange(Ns), np.arange(Ns), indexing="ij"), -1).reshape(-1, 3)
    arg = q.astype(np.float64) * (box / np.float64(Ns)) + D.astype(np.float64) + D2.astype(np.float64)
    pos = periodic_wrap(arg.astype(np.float32), box)
    vel = np.zeros_like(pos) if use_cola else (dDdy.astype(np.float32) + dD2dy.astype(np.float32)).astype(np.float32)
    ids = (q[:, 0].astype(np.uint64) * Ns + q[:, 1].astype(np.uint64)) * Ns + q[:, 2].astype(np.uint64)
    return pos, vel, ids


# ----------------------------------------------------------------------------- massive neutrinos (auxPM.c:383-420)

def nu_add(P3D, cdelta_cdm, nufac_by_k2, cdmfac, nmesh):
    """P3D = cdmfac * P3D + nufac(|k|) * cdelta_cdm for all modes but (0,0,0); nufac_by_k2[m] at m = |d|^2 is
    OmegaNu / Omega * Nmesh^3 * T_nu(k, a) / T_cb(k, 1) (auxPM.c:394, 407)."""
    RK = _rk(nmesh).astype(np.int64)
    nf = np.asarray(nufac_by_k2, dtype=np.float64)[RK]
    out = cdmfac * P3D + nf * cdelta_cdm
    out[0, 0, 0] = P3D[0, 0, 0]
    return out


# ----------------------------------------------------------------------------- redshift-space multipoles (compute_pofk.c:280-753)

def ptomesh_rsd_deposit(pos, V, axis, vnorm, nmesh, nsample, box, grid_dtype=np.float64):
    """PtoMesh_RSD (compute_pofk.c:280-393), single task.  V = the line-of-sight velocity per particle as the
    reference forms it (double): Vel[axis] (+ the COLA LPT velocity).  axis 1 = y (y and z swap roles), 2 = z."""
    N = nmesh
    nzp = 2 * (N // 2 + 1)
    scale = np.float64(N) / np.float64(box)
    wpar = (np.float64(N) / np.float64(nsample)) ** 3
    X = pos[:, 0].astype(np.float64) * scale
    if axis == 1:
        Y = pos[:, 2].astype(np.float64) * scale
        Z = pos[:, 1].astype(np.float64) * scale
    else:
        Y = pos[:, 1].astype(np.float64) * scale
        Z = pos[:, 2].astype(np.float64) * scale
    Z = Z + np.asarray(V, dtype=np.float64) * vnorm
    Z = np.where(Z >= N, Z - N, Z)
    Z = np.where(Z < 0, Z + N, Z)
    IX, IY, IZ = X.astype(np.uint32).astype(np.int64), Y.astype(np.uint32).astype(np.int64), Z.astype(np.uint32).astype(np.int64)
    DX, DY, DZ = X - IX, Y - IY, Z - IZ
    TX, TY, TZ = 1.0 - DX, 1.0 - DY, 1.0 - DZ
    DY = DY * wpar
    TY = TY * wpar
    IY[IY >= N] = 0
    IZ[IZ >= N] = 0
    IXn, IYn, IZn = IX + 1, IY + 1, IZ + 1
    IYn[IYn >= N] = 0
    IZn[IZn >= N] = 0
    size = (N + 1) * N * nzp
    grid = np.zeros(size, dtype=np.float64)
    for ix, wx in ((IX, TX), (IXn, DX)):
        for iy, wy in ((IY, TY), (IYn, DY)):
            for iz, wz in ((IZ, TZ), (IZn, DZ)):
                grid += np.bincount((ix * N + iy) * nzp + iz, weights=wx * wy * wz, minlength=size)
    grid = (grid - 1.0).astype(grid_dtype).reshape(N + 1, N, nzp)
    grid[0] += (grid[N] + grid_dtype(1.0)).astype(grid_dtype)
    return np.ascontiguousarray(grid[:N])


def bin_up_rsd_sums(dens_k, nmesh, box, nbins, bintype, subtract_shotnoise, kmin_hmpc, kmax_hmpc):
    """The five per-bin sums of bin_up_RSD_power_spectrum before normalisation (compute_pofk.c:568-690):
    returns (pofk0, pofk2, pofk4, n, k) arrays and the adjusted (nbins, bintype, shot, kmin, kmax)."""
    N = nmesh
    nbins, bintype, shot, kmin, kmax = adjust_pofk_parameters(N, box, nbins, bintype, subtract_shotnoise, kmin_hmpc, kmax_hmpc)
    i = np.arange(N)
    dxy = np.where(i > N // 2, N - i, i)
    dz = np.arange(N // 2 + 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        gxy = np.where(dxy == 0, 1.0, np.sin((PI * dxy) / float(N)) / ((PI * dxy) / float(N)))
        gz = np.where(dz == 0, 1.0, np.sin((PI * dz) / float(N)) / ((PI * dz) / float(N)))
    gz[N // 2] = 2.0 / PI
    m = (dxy[:, None, None] ** 2 + dxy[None, :, None] ** 2 + dz[None, None, :] ** 2).astype(np.float64)
    kmag = np.sqrt(m)
    with np.errstate(invalid="ignore", divide="ignore"):
        mu2 = (dz[None, None, :] ** 2).astype(np.float64) / kmag / kmag
    mu2[0, 0, 0] = 0.0
    nk = pofk_bin_index(kmag, kmin, kmax, nbins, bintype)
    corr = 1.0 / (gxy[:, None, None] * gxy[None, :, None] * gz[None, None, :]) ** 4.0 * (1.0 / np.float64(N) ** 6)
    p = (dens_k.real * dens_k.real + dens_k.imag * dens_k.imag) * corr
    w = np.full(N // 2 + 1, 2.0)
    w[0] = 1.0
    w[N // 2] = 1.0
    w = np.broadcast_to(w[None, None, :], p.shape)
    ok = (nk >= 0) & (nk < nbins)
    mu2 = np.broadcast_to(mu2, p.shape)

    def acc(v):
        return np.bincount(nk[ok], weights=v[ok], minlength=nbins)
    return (acc(w * p), acc(w * p * mu2), acc(w * p * mu2 * mu2), acc(w), acc(w * kmag)), (nbins, bintype, shot, kmin, kmax)


def rsd_multipoles(sums, cfg, box, nsample):
    """Normalisation, Legendre combinations, shot noise on all three, bin-centre k (compute_pofk.c:699-725)."""
    p0s, p2s, p4s, n, ks = sums
    nbins, bintype, shot, kmin, kmax = cfg
    good = n > 0
    P0, P2, P4, k = np.zeros(nbins), np.zeros(nbins), np.zeros(nbins), np.zeros(nbins)
    b3 = box ** 3
    p0 = p0s[good] / n[good] * b3
    p2 = p2s[good] / n[good] * b3
    p4 = p4s[good] / n[good] * b3
    p4 = 9.0 * (35.0 * p4 - 30.0 * p2 + 3.0 * p0) / 8.0
    p2 = 5.0 * (3.0 * p2 - 1.0 * p0) / 2.0
    if shot:
        sn = (box / float(nsample)) ** 3
        p0, p2, p4 = p0 - sn, p2 - sn, p4 - sn
    idx = np.arange(nbins)[good]
    if bintype == 0:
        kk = kmin + (kmax - kmin) / float(nbins) * idx
    else:
        kk = np.exp(np.log(kmin) + np.log(kmax / kmin) / float(nbins) * idx)
    P0[good], P2[good], P4[good], k[good] = p0, p2, p4, kk * 2.0 * np.pi / box
    return n, k, P0, P2, P4


def compute_rsd_powerspectrum(pos, Vy, Vz, vnorm, nmesh, nsample, box, pofk, grid_dtype=np.float64):
    """compute_RSD_powerspectrum (compute_pofk.c:403-512): lines of sight y then z.  Returns the two per-axis
    results (n, k, P0, P2, P4) and the raw sums."""
    out = {}
    for name, axis, V in (("y", 1, Vy), ("z", 2, Vz)):
        dens = ptomesh_rsd_deposit(pos, V, axis, vnorm, nmesh, nsample, box, grid_dtype)
        dk = r2c(dens, nmesh)
        sums, cfg = bin_up_rsd_sums(dk, nmesh, box, **pofk)
        out[name] = rsd_multipoles(sums, cfg, box, nsample)
        out[name + "_sums"] = sums
    return out


def rsd_velocity(vel, D, D2, axis, use_cola, dDdy, dD2dy, scale_dependent=False):
    """The line-of-sight velocity of PtoMesh_RSD (compute_pofk.c:320-327): SCALEDEPENDENT: D, D2 are P.dDdy, P.dD2dy
    and the sum is float arithmetic; else Vel + (D dDdy + D2 dD2dy) in double."""
    v = vel[:, axis]
    if not use_cola:
        return v.astype(np.float64)
    if scale_dependent:
        return (v + (D[:, axis] + D2[:, axis]).astype(np.float32)).astype(np.float32).astype(np.float64)
    return v.astype(np.float64) + (D[:, axis].astype(np.float64) * dDdy + D2[:, axis].astype(np.float64) * dD2dy)


# ----------------------------------------------------------------------------- SimplePofk (post-processing tool)

def simple_pofk(pos, ngrid, box, scheme="CIC", subtract_shotnoise=False):
    """SimplePofk/main.cpp restated (the reference's stand-alone P(k) estimator; the only user of NGP and TSC):
    positions as GADGET floats, x = double(pos_float / boxsize) (main.cpp:259-261); assignment add_to_grid_NGP / _CIC /
    _TSC (59-228) of counts; main()'s normalisation to the density contrast (513-533: grid /= N^3 / Npart, the mean
    of that, grid = grid / mean - 1); complex 3-D transform; per mode |d_k|^2 / N^6 / window^2 with window =
    prod_a sinc(pi k_a / N) to the power 1 / 2 / 3 (324-340, 393); bin int(|k| + 0.5) for 0 < bin < N (391); mean per
    bin, optional shot noise 1 / Npart (420-424).
    TSC reproduces the tool as written, including its three `izneighN` where the weight PZ belongs to `izneighP`
    (main.cpp:189, 201, 213: the "010", "110", "210" lines).  Returns (pofk[N], nmodes[N]); the tool prints
    k = (2 i + 1) pi / boxsize and pofk[i] * boxsize^3 for 1 <= i <= N / 2 (436-444).
    PINNED to the tool itself: oracle/_ref/simplepofk_{NGP,CIC,TSC} = SimplePofk/main.cpp compiled unmodified (g++ -fopenmp,
    the FFT stand-in's complex plan) run on a GADGET file of these positions (tests/test_simple_pofk.py), to the six digits
    the tool prints."""
    N = int(ngrid)
    p = np.asarray(pos, dtype=np.float32)
    x = (p.astype(np.float64) / np.float64(box))
    X = x * N
    I = X.astype(np.int64)
    Dd = X - I
    I = np.where(I >= N, I - N, I)
    grid = np.zeros((N, N, N))                      # grid[ix, iy, iz]; the tool stores ix fastest, P(k) does not care

    def add(ix, iy, iz, w):
        np.add.at(grid, (ix, iy, iz), w)
    ix, iy, iz = I[:, 0], I[:, 1], I[:, 2]
    if scheme == "NGP":
        add(ix, iy, iz, np.ones(len(p)))
        power = 1
    elif scheme == "CIC":
        nb = [np.where(I[:, a] + 1 >= N, I[:, a] + 1 - N, I[:, a] + 1) for a in range(3)]
        T, Dw = 1.0 - Dd, Dd
        for ax, wx in ((ix, T[:, 0]), (nb[0], Dw[:, 0])):
            for ay, wy in ((iy, T[:, 1]), (nb[1], Dw[:, 1])):
                for az, wz in ((iz, T[:, 2]), (nb[2], Dw[:, 2])):
                    add(ax, ay, az, wx * wy * wz)
        power = 2
    elif scheme == "TSC":
        Tw = 0.75 - Dd * Dd
        Nw = 0.5 * (0.5 + Dd) * (0.5 + Dd)
        Pw = 0.5 * (0.5 - Dd) * (0.5 - Dd)
        nN = [np.where(I[:, a] + 1 >= N, I[:, a] + 1 - N, I[:, a] + 1) for a in range(3)]
        nP = [np.where(I[:, a] - 1 < 0, I[:, a] - 1 + N, I[:, a] - 1) for a in range(3)]
        cx = ((nP[0], Pw[:, 0]), (ix, Tw[:, 0]), (nN[0], Nw[:, 0]))
        cy = ((nP[1], Pw[:, 1], "P"), (iy, Tw[:, 1], "T"), (nN[1], Nw[:, 1], "N"))
        for ax, wx in cx:
            for ay, wy, ytag in cy:
                # the z triple: previous, this, next -- except that on the "this y" lines the previous-z weight lands on next z
                zP = nN[2] if ytag == "T" else nP[2]
                add(ax, ay, zP, wx * wy * Pw[:, 2])
                add(ax, ay, iz, wx * wy * Tw[:, 2])
                add(ax, ay, nN[2], wx * wy * Nw[:, 2])
        power = 3
    else:
        raise ValueError(scheme)
    grid /= float(N) ** 3 / float(len(p))           # main.cpp:515-521 (sic: divided, the mean below undoes it)
    grid = grid / grid.mean() - 1.0                 # 523-528
    dk = np.fft.fftn(grid)
    k1 = np.arange(N)
    kk = np.where(k1 < N // 2, k1, k1 - N)
    ii, jj, ll = np.meshgrid(kk, kk, kk, indexing="ij")
    kind = (np.sqrt((ii * ii + jj * jj + ll * ll).astype(np.float64)) + 0.5).astype(np.int64)
    fac = np.pi / N

    def sinc(k):
        a = k * fac
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(k != 0, np.sin(a) / np.where(k != 0, a, 1.0), 1.0)
    w = (sinc(ii) * sinc(jj) * sinc(ll)) ** power
    val = (dk.real ** 2 + dk.imag ** 2) * (1.0 / float(N) ** 3) ** 2 / (w * w)
    sel = (kind < N) & (kind > 0)
    pofk = np.bincount(kind[sel], weights=val[sel], minlength=N).astype(np.float64)
    nmodes = np.bincount(kind[sel], minlength=N).astype(np.float64)
    good = nmodes > 0
    pofk[good] /= nmodes[good]
    if subtract_shotnoise:
        pofk -= 1.0 / float(len(p))
    return pofk, nmodes


# ----------------------------------------------------------------------------- READICFROMFILE

def readic_delta_k(pos01_files, nmesh, nsample, normfac, rescale_by_k2, grid_dtype=np.float64):
    """ReadFilesMakeDisplacementField + AssignDisplacementField restated for one task (readICfromfile.c:133-215,
    533-778): CIC of the external particles (coordinates in [0, 1), X = pos * Nmesh, W = (Nmesh/Nsample)^3) onto a grid
    that starts at -1, the ghost plane folded onto plane 0 (630-637), r2c, density *= normfac (642-646), sharp-k filter
    above Nsample/2 when Nmesh > Nsample (648-680), and cdelta_cdm = P3D * grid_corr * rescale_fac with the CIC window
    grid_corr = prod_a (sin(pi d_a / N) / (pi d_a / N))^-2 (724-750).  Returns delta_k [N][N][N/2+1]; mode 0 is 0.
    PINNED to the unmodified reference built with -DREADICFROMFILE -DSCALEDEPENDENT (oracle/_ref/libmgpicola_ref_fofr_ric.so)
    reading GADGET files written by the test: its cdelta_cdm to 1e-12, with Nmesh = and > Nsample and both settings of
    input_sigma8_is_for_lcdm (tests/test_readic_oracle.py)."""
    N = int(nmesh)
    W = (float(nmesh) / float(nsample)) ** 3
    dens = np.full((N + 1, N, N), -1.0)
    for pos in pos01_files:
        p = np.asarray(pos, dtype=np.float32).astype(np.float64) * float(N)
        I = p.astype(np.uint32).astype(np.int64)
        D = p - I
        T = 1.0 - D
        ix = I[:, 0]
        iy = np.where(I[:, 1] >= N, 0, I[:, 1])
        iz = np.where(I[:, 2] >= N, 0, I[:, 2])
        iy1 = np.where(iy + 1 >= N, 0, iy + 1)
        iz1 = np.where(iz + 1 >= N, 0, iz + 1)
        Dy, Ty = D[:, 1] * W, T[:, 1] * W
        for ax, wx in ((ix, T[:, 0]), (ix + 1, D[:, 0])):
            for ay, wy in ((iy, Ty), (iy1, Dy)):
                for az, wz in ((iz, T[:, 2]), (iz1, D[:, 2])):
                    np.add.at(dens, (ax, ay, az), wx * wy * wz)
    dens[0] += dens[N] + 1.0                                   # the slice from the left task: density += temp + 1
    g = dens[:N].astype(grid_dtype)
    dk = np.fft.rfftn(g.astype(np.float64)).astype(np.complex64 if grid_dtype == np.float32 else np.complex128)
    dk = (dk * normfac).astype(dk.dtype)
    k1 = np.arange(N)
    d0 = np.where(k1 > N // 2, k1 - N, k1)
    dd0, dd1, dd2 = np.meshgrid(d0, d0, np.arange(N // 2 + 1), indexing="ij")
    m = dd0 * dd0 + dd1 * dd1 + dd2 * dd2
    if nmesh > nsample:
        dk = np.where(np.sqrt(m.astype(np.float64)) > float(nsample // 2), 0.0, dk)

    def sinc(d):
        a = PI * d / float(N)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(d != 0, np.sin(a) / np.where(d != 0, a, 1.0), 1.0)
    gc = (1.0 / (sinc(dd0) * sinc(dd1) * sinc(dd2))) ** 2
    out = (dk.astype(np.complex128) * gc) * np.asarray(rescale_by_k2, dtype=np.float64)[m]
    out[0, 0, 0] = 0.0
    return out


# ----------------------------------------------------------------------------- lightcone (-DLIGHTCONE)

def cspline_coeffs(x, y):
    """Coefficient array of gsl_interp_cspline (natural cubic spline; GSL interpolation/cspline.c, a third-party
    dependency absent from the reference tree -- its published construction): c[0] = c[n-1] = 0, the interior from the
    symmetric tridiagonal system h_i c_i + 2 (h_i + h_{i+1}) c_{i+1} + h_{i+1} c_{i+2} = 3 (dy_{i+1}/h_{i+1} - dy_i/h_i).
    Call sites: lightcone.c:349-357."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    n = x.size
    c = np.zeros(n)
    if n < 3:
        return c
    h = np.diff(x)
    dy = np.diff(y)
    off = h[1:].copy()
    diag = 2.0 * (h[1:] + h[:-1])
    g = 3.0 * (dy[1:] / h[1:] - dy[:-1] / h[:-1])
    m = n - 2
    for i in range(1, m):
        w = off[i - 1] / diag[i - 1]
        diag[i] -= w * off[i - 1]
        g[i] -= w * g[i - 1]
    c[m] = g[m - 1] / diag[m - 1]
    for i in range(m - 2, -1, -1):
        c[i + 1] = (g[i] - off[i] * c[i + 2]) / diag[i]
    return c


def cspline_eval(x, y, c, v):
    """gsl_spline_eval of the cubic spline (cspline.c: b = dy/dx - dx (c[i+1] + 2 c[i]) / 3, d = (c[i+1] - c[i]) / (3 dx),
    y = y[i] + t (b + t (c[i] + t d))); interval x[i] <= v < x[i+1], the last one for v = x[n-1]."""
    v = np.asarray(v, np.float64)
    i = np.clip(np.searchsorted(x, v, side="right") - 1, 0, x.size - 2)
    dx = x[i + 1] - x[i]
    dy = y[i + 1] - y[i]
    b = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0
    d = (c[i + 1] - c[i]) / (3.0 * dx)
    t = v - x[i]
    return y[i] + t * (b + t * (c[i] + t * d))


def drift_lightcone(pos, vel, D, D2, sumxyz, box, use_cola, A, AFF, dyyy, da1, da2, dv1, dv2, rcomov_old, rcomov_new,
                    origin, reps, al_tab, da1_tab, da2_tab, dyyy_tab, lengthfac, velfac_times_fac, boundary=20.0):
    """The particle loop of Drift_Lightcone (lightcone.c:392-471) for the replicates `reps` ([nrep][3] offsets with
    repflag == 0, in the order of the triple loop 411-413).  Returns (new positions float32, [rows of replicate r as
    float32 [count][6], in particle order], exceeded) with exceeded = True where the reference calls FatalError (403-407)."""
    P = pos.astype(np.float64)
    V = vel.astype(np.float64) - np.asarray(sumxyz, np.float64)[None, :]
    Dd = D.astype(np.float64)
    D2d = D2.astype(np.float64)
    uc = float(use_cola)
    dpos = V * dyyy + uc * (Dd * da1 + D2d * da2)                               # 398-400
    exceeded = bool(np.any(dpos > boundary))                                    # 403
    al_tab = np.asarray(al_tab, np.float64)
    tabs = [np.asarray(t, np.float64) for t in (da1_tab, da2_tab, dyyy_tab)]
    cs = [cspline_coeffs(al_tab, t) for t in tabs]
    org = np.asarray(origin, np.float64)
    rows = []
    for (i, j, k) in np.asarray(reps, np.int64).reshape(-1, 3):
        shift = np.array([i * box, j * box, k * box], np.float64)
        X = P - org[None, :] + shift[None, :]                                   # 420-422
        ro2 = X[:, 0] * X[:, 0] + X[:, 1] * X[:, 1] + X[:, 2] * X[:, 2]
        Xn = X + dpos
        rn2 = Xn[:, 0] * Xn[:, 0] + Xn[:, 1] * Xn[:, 1] + Xn[:, 2] * Xn[:, 2]
        sel = np.nonzero((ro2 <= rcomov_old * rcomov_old) & (rn2 > rcomov_new * rcomov_new))[0]   # 425, 434
        ro, rn = np.sqrt(ro2[sel]), np.sqrt(rn2[sel])
        AL = A + (AFF - A) * ((rcomov_old - ro) / ((rn - ro) - (rcomov_new - rcomov_old)))        # 440
        t1, t2, ty = (cspline_eval(al_tab, tabs[q], cs[q], AL) for q in range(3))
        out = np.empty((sel.size, 6), np.float32)
        x = P[sel] + V[sel] * ty[:, None] + uc * (Dd[sel] * t1[:, None] + D2d[sel] * t2[:, None]) + shift[None, :]   # 447-449
        out[:, :3] = (lengthfac * x).astype(np.float32)
        out[:, 3:] = (velfac_times_fac * (V[sel] + (Dd[sel] * dv1 + D2d[sel] * dv2) * uc)).astype(np.float32)     # 451-453
        rows.append(out)
    newpos = periodic_wrap((P + dpos).astype(np.float32), box)                  # 468-470
    return newpos, rows, exceeded


# ----------------------------------------------------------------------------- FoF halo finder (-DMATCHMAKER_HALOFINDER)

FOF_HALO_DTYPE = np.dtype([("np", np.int32), ("m_halo", np.float32), ("x_avg", np.float32, 3), ("x_rms", np.float32, 3),
                           ("v_avg", np.float32, 3), ("v_rms", np.float32, 3), ("lam", np.float32, 3), ("b", np.float32),
                           ("c", np.float32), ("ea", np.float32, 3), ("eb", np.float32, 3), ("ec", np.float32, 3)])   # mm_common.h:115-128


def fof_translate(pos, vel, D, D2, norm_pos, norm_vel, use_cola, dDdy, dD2dy, scale_dependent=False):
    """picola_to_matchmaker_particles, mm_main.c:294-338: MatchMaker's float positions (before the x offset) and
    velocities.  scale_dependent: D / D2 are then P.dDdy / P.dD2dy (float sum, mm_main.c:319)."""
    npf, nvf = np.float32(norm_pos), np.float32(norm_vel)
    x = (npf * pos.astype(np.float32)).astype(np.float32)
    if not use_cola:
        v = (nvf * vel.astype(np.float32)).astype(np.float32)
    elif scale_dependent:
        v = (nvf * (vel.astype(np.float32) + (D.astype(np.float32) + D2.astype(np.float32)))).astype(np.float32)
    else:
        v = (np.float64(nvf) * (vel.astype(np.float64) + (D.astype(np.float64) * dDdy + D2.astype(np.float64) * dD2dy))).astype(np.float32)
    return x, v


def fof_link_pairs(x, dfof, lbox):
    """All pairs (i < j) that get_neighbors (mm_fof.c:112-186) links: dx = |x0 - x1| (NOT periodic: slab coordinates),
    dy, dz periodic (d > L/2 -> L - d), d2 = dx dx + dy dy + dz dz <= Dfof^2, everything in float."""
    from scipy.spatial import cKDTree
    dfof = np.float32(dfof)
    lb, lh = np.float32(lbox), np.float32(lbox / 2)
    xd = x.astype(np.float64).copy()
    L = float(lb)
    xd[:, 1:] = np.mod(xd[:, 1:], L)
    xd[:, 1:][xd[:, 1:] >= L] = 0.0
    span = float(xd[:, 0].max() - xd[:, 0].min()) + 4.0 * float(dfof) + 1.0
    xd[:, 0] -= xd[:, 0].min()
    tree = cKDTree(xd, boxsize=[span * 4.0, L, L])                  # x: a period no pair can wrap around
    cand = tree.query_pairs(float(dfof) * 1.0001 + 1e-6, output_type="ndarray")
    i, j = cand[:, 0], cand[:, 1]
    dx = np.abs(x[i, 0] - x[j, 0]).astype(np.float32)
    dy = np.abs(x[i, 1] - x[j, 1]).astype(np.float32)
    dz = np.abs(x[i, 2] - x[j, 2]).astype(np.float32)
    dy = np.where(dy > lh, lb - dy, dy).astype(np.float32)
    dz = np.where(dz > lh, lb - dz, dz).astype(np.float32)
    d2 = ((dx * dx + dy * dy).astype(np.float32) + dz * dz).astype(np.float32)
    ok = (dx <= dfof) & (d2 <= np.float32(dfof * dfof))
    return i[ok], j[ok]


def fof_halo_properties(x, v, ids, boxsize, mass_particle, x_offset):
    """get_halos (mm_fof.c:468-611) for one group: members `ids` in increasing index order."""
    L = float(boxsize)
    n = ids.size
    h = np.zeros((), FOF_HALO_DTYPE)
    h["np"] = n
    h["m_halo"] = n * mass_particle
    xs, vs = x[ids].astype(np.float64), v[ids].astype(np.float64)
    x_sum = np.zeros(3)
    for j in range(n):                                              # running centre of mass decides the image (490-511)
        for ax in range(3):
            xx = xs[j, ax]
            if j > 0:
                cm = x_sum[ax] / j
                if 2 * abs(xx - cm) > L:
                    xx = xx - L if 2 * xx > L else xx + L
            x_sum[ax] += xx
    v_sum = np.zeros(3)
    for j in range(n):
        v_sum += vs[j]
    x_avg = (x_sum / n).astype(np.float32)                          # stored as float (512-515)
    v_avg = (v_sum / n).astype(np.float32)
    xa, va = x_avg.astype(np.float64), v_avg.astype(np.float64)
    xw = xs.copy()
    far = 2 * np.abs(xw - xa[None, :]) > L                          # 527-536
    xw = np.where(far, np.where(2 * xw > L, xw - L, xw + L), xw)
    dx, dv = xw - xa[None, :], vs - va[None, :]
    x_rms, v_rms, lam, inertia = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros((3, 3))
    for j in range(n):                                              # 539-555, in member order
        x_rms += dx[j] * dx[j]
        v_rms += dv[j] * dv[j]
        inertia += np.outer(dx[j], dx[j])
        lam[0] += dx[j, 1] * dv[j, 2] - dx[j, 2] * dv[j, 1]
        lam[1] += dx[j, 2] * dv[j, 0] - dx[j, 0] * dv[j, 2]
        lam[2] += dx[j, 0] * dv[j, 1] - dx[j, 1] * dv[j, 0]
    w, e = np.linalg.eigh(inertia)                                  # gsl_eigen_symmv + descending sort (557-568)
    o = np.argsort(-w, kind="stable")
    w, e = w[o], e[:, o]
    if w[0] <= 0:
        h["b"] = h["c"] = 0
    else:
        h["b"], h["c"] = w[1] / w[0], w[2] / w[0]
        h["ea"], h["eb"], h["ec"] = e[:, 0], e[:, 1], e[:, 2]
    h["x_rms"] = np.sqrt(x_rms / n)
    h["v_rms"] = np.sqrt(v_rms / n)
    h["lam"] = lam
    xf = x_avg.copy()                                               # wrap the centre of mass (598-603), float arithmetic
    for ax in range(3):
        if xf[ax] < 0:
            xf[ax] = np.float32(np.float64(xf[ax]) + L)
        elif xf[ax] >= L:
            xf[ax] = np.float32(np.float64(xf[ax]) - L)
    xf[0] = np.float32(xf[0] + np.float32(x_offset))
    h["x_avg"] = xf
    h["v_avg"] = v_avg
    return h


def fof_halos(tasks, norm_pos, norm_vel, boxsize, dx_extra, b_fof, np_min, mass_part, n_part_1d, use_cola=1, dDdy=0.0,
              dD2dy=0.0, scale_dependent=False):
    """MatchMaker (mm_main.c:129-385, mm_fof.c) on NTask = len(tasks) tasks: tasks[t] = dict(pos, vel, D, D2,
    local_p_start) holds task t's particles (D / D2 = P.dDdy / P.dD2dy when scale_dependent).  Returns per task the
    FoFHalo records sorted by np (descending), and per task a dict of intermediate results for the tests."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    nt = len(tasks)
    L = float(boxsize)
    ipd = np.float32(L / np.float64(float(n_part_1d) ** 3) ** (1.0 / 3.0))        # init_fof, mm_fof.c:83
    dfof = np.float32(np.float64(ipd) * b_fof)
    mass_particle = 1.0e10 * mass_part                                             # MATCHMAKER_MASS_FACTOR
    x_off = [np.float32(t["local_p_start"] * (L / float(n_part_1d))) for t in tasks]                  # mm_main.c:227
    st = []
    for r, t in enumerate(tasks):
        xo_right = x_off[(r + 1) % nt]
        dxd = np.float32(xo_right - x_off[r])                                       # 236-237
        if dxd < 0:
            dxd = np.float32(np.float64(dxd) + L)
        pos = t["pos"]
        x1 = (np.float64(norm_pos) * pos[:, 0].astype(np.float64) - np.float64(x_off[r])).astype(np.float32)   # 266-271
        n_toleft = int(np.count_nonzero(x1.astype(np.float64) <= dx_extra))
        x, v = fof_translate(pos, t["vel"], t["D"], t["D2"], norm_pos, norm_vel, use_cola, dDdy, dD2dy, scale_dependent)
        x[:, 0] = x[:, 0] - x_off[r]                                                # 334
        order = np.argsort(x[:, 0], kind="stable")                                  # qsort by x (360)
        st.append(dict(x=x[order], v=v[order], order=order, n_dom=pos.shape[0], n_toleft=n_toleft, dx_domain=dxd))
    for r in range(nt):                                                             # buffer from the right neighbour (363-373)
        right = st[(r + 1) % nt]
        bx = right["x"][:right["n_toleft"]].copy()
        bx[:, 0] = bx[:, 0] + st[r]["dx_domain"]
        st[r]["xa"] = np.concatenate([st[r]["x"], bx])
        st[r]["va"] = np.concatenate([st[r]["v"], right["v"][:right["n_toleft"]]])
    for r in range(nt):                                                             # assign_particles_to_fof (354-400)
        s = st[r]
        n = s["xa"].shape[0]
        i, j = fof_link_pairs(s["xa"], dfof, L)
        ncomp, lab = connected_components(coo_matrix((np.ones(i.size, np.int8), (i, j)), shape=(n, n)), directed=False)
        size = np.bincount(lab, minlength=ncomp)
        has_dom = np.bincount(lab[:s["n_dom"]], minlength=ncomp) > 0               # groups are seeded by in-domain particles only
        s["lab"], s["in_group"] = lab, (size[lab] >= 2) & has_dom[lab]
    out, info = [], []
    for r in range(nt):
        s, left = st[r], st[(r - 1) % nt]
        member = s["in_group"].copy()
        back = left["in_group"][left["n_dom"]:]                                     # my first n_toleft particles as the left task saw them (425-437)
        member[:s["n_toleft"]] &= ~back
        lab = s["lab"]
        cnt = np.bincount(lab[member], minlength=lab.max() + 1)
        halos = []
        for g in np.nonzero(cnt >= np_min)[0]:
            ids = np.nonzero(member & (lab == g))[0]
            halos.append(fof_halo_properties(s["xa"], s["va"], ids, L, mass_particle, x_off[r]))
        h = np.array(halos, FOF_HALO_DTYPE) if halos else np.zeros(0, FOF_HALO_DTYPE)
        h = h[np.argsort(-h["np"], kind="stable")]
        out.append(h)
        info.append(dict(n_groups=int(np.count_nonzero(cnt > 0)), n_toleft=s["n_toleft"], n_fromright=s["xa"].shape[0] - s["n_dom"],
                         dfof=float(dfof), group_sizes=np.sort(cnt[cnt > 0])[::-1]))
    return out, info
