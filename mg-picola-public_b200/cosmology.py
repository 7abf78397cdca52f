"""Host-side scalars of a COLA run (what stays in C in the real driver).

In the drop-in deployment main.c / cosmo.c / user_defined_functions.h keep computing these numbers
and hand them to the CUDA library through include/mgpicola.h.  bench.py and the Python run driver
(`run.py`) have no C driver above them, so this module restates the *scalar* set-up they need:

  * LCDM background and first / second order growth factors: the ODE system of cosmo.c:238-262
    (`ode_growth_DLCDM`) integrated in x = ln a from z = max(200, z_init) exactly as
    `solve_for_growth_factors` (cosmo.c:409-530) stores it: D, dD/dy = D' Q/a, ddD/ddy = 1.5 Omega a D,
    D2, dD2/dy, ddD2/ddy = 1.5 Omega a (D2 - D^2), all normalised at a = 1 (cosmo.c:318-403).
  * the COLA step integrals Sq / Sphi of cosmo.c:1114-1160 (StdDA = 0, fullT = 1, nLPT = -2.5).
  * the time-step schedule of main.c:394-449, 598-602 (stepDistr = 0: linear in a).
  * per-step modified-gravity scalars (user_defined_functions.h:453-455, 462-470, 490-498, 587-595,
    725-737, 757-766; mg.h:27-28, 79-80).

Nothing here runs on the GPU and nothing here is on the per-particle / per-cell path.
"""
import numpy as np
from scipy.integrate import quad, solve_ivp
from scipy.interpolate import CubicSpline

INVERSE_H0_MPCH = 2997.92458   # vars.h:64
N_LPT = -2.5                   # main.c:82-85


class LCDM:
    def __init__(self, omega, z_init=9.0, npts=1000):
        self.omega = float(omega)
        zini = max(200.0, z_init)                      # cosmo.c:51
        xini, xend = np.log(1.0 / (1.0 + zini)), np.log(1.0 / (1.0 - 0.5))   # cosmo.c:52
        x = np.linspace(xini, xend, npts)
        om = self.omega

        def rhs(xx, y):
            a = np.exp(xx)
            H = self.hubble(a)
            dH = self.dhubbleda(a)
            beta = 1.5 * om / (a * a * a * H * H)
            alpha = 2.0 + a * dH / H
            return [y[1], -alpha * y[1] + beta * y[0], y[3], -alpha * y[3] + beta * (y[2] - y[0] * y[0])]

        sol = solve_ivp(rhs, (xini, xend), [1.0, 1.0, -3.0 / 7.0, -6.0 / 7.0], t_eval=x, rtol=1e-10, atol=1e-12,
                        method="DOP853")
        a = np.exp(x)
        Q = self.qfactor(a)
        D, q, D2, q2 = sol.y
        self._D = CubicSpline(x, D, bc_type="natural")
        self._dD = CubicSpline(x, q * Q / a, bc_type="natural")
        self._ddD = CubicSpline(x, 1.5 * om * a * D, bc_type="natural")
        self._D2 = CubicSpline(x, D2, bc_type="natural")
        self._dD2 = CubicSpline(x, q2 * Q / a, bc_type="natural")
        self._ddD2 = CubicSpline(x, 1.5 * om * a * (D2 - D * D), bc_type="natural")
        self._n1 = float(self._D(0.0))
        self._n2 = float(self._D2(0.0))

    # user_defined_functions.h:413-447 (LCDM background)
    def hubble(self, a):
        return np.sqrt(self.omega / (a * a * a) + 1.0 - self.omega)

    def dhubbleda(self, a):
        return 1.0 / (2.0 * self.hubble(a)) * (-3.0 * self.omega / (a * a * a * a))

    def qfactor(self, a):            # cosmo.c:197-199
        return self.hubble(a) * a * a * a

    def growth_D(self, a): return float(self._D(np.log(a))) / self._n1
    def growth_dDdy(self, a): return float(self._dD(np.log(a))) / self._n1
    def growth_ddDddy(self, a): return float(self._ddD(np.log(a))) / self._n1
    def growth_D2(self, a): return float(self._D2(np.log(a))) / self._n2
    def growth_dD2dy(self, a): return float(self._dD2(np.log(a))) / self._n2
    def growth_ddD2ddy(self, a): return float(self._ddD2(np.log(a))) / self._n2

    # cosmo.c:1114-1160 with fullT = 1
    def Sq(self, ai, af, aref):
        res, _ = quad(lambda a: a ** N_LPT / self.qfactor(a), ai, af, epsrel=1e-8)
        return res / aref ** N_LPT

    def Sphi(self, ai, af, aref):
        return (af ** N_LPT - ai ** N_LPT) * aref / self.qfactor(aref) / (N_LPT * aref ** (N_LPT - 1.0))


def schedule(z_init, outputs):
    """The (A, AI, AF, AFF, is_output_step, is_final) sequence of main.c:394-449, 598-602 for
    stepDistr = 0.  outputs = [(z_out, nsteps), ...]."""
    A = 1.0 / (1.0 + z_init)
    AI = A
    seq = []
    nout = len(outputs)
    for i in range(nout + 1):
        if i == nout:
            nsteps, da = 1, 0.0
        else:
            ao = 1.0 / (1.0 + outputs[i][0])
            nsteps = outputs[i][1]
            da = (ao - A) / float(nsteps)
        for ts in range(nsteps):
            out_step = (ts == 0 and i != 0)
            AF = A if out_step else A + da * 0.5
            AFF = A + da
            seq.append(dict(A=A, AI=AI, AF=AF, AFF=AFF, da=da, output=out_step, final=(i == nout), iout=i))
            if i == nout:
                break
            if out_step:
                AI2, AF2 = A, A + da * 0.5      # second kick of an output step (main.c:559-569)
                seq[-1].update(AI2=AI2, AF2=AF2)
                AF = AF2
            A, AI = AFF, AF
    return seq


def fofr_step_scalars(a, omega, box, fofr0, nfofr):
    """phi_crit (udf:731-734), coupling 2 beta^2 = 1/3 (udf:462-470, 587-591), massterm2 (mg.h:79-80,
    udf:490-498)."""
    phicrit = 1.5 * fofr0 * ((omega + 4.0 * (1.0 - omega)) / (omega / (a * a * a) + 4.0 * (1.0 - omega))) ** (nfofr + 1.0)
    coupling = 2.0 * (1.0 / np.sqrt(6.0)) ** 2
    fac = omega / (a * a * a) + 4.0 * (1.0 - omega)
    fac0 = omega + 4.0 * (1.0 - omega)
    mass2 = fac0 * (fac / fac0) ** (nfofr + 2.0) / ((1.0 + nfofr) * fofr0)
    massterm2 = a * a * mass2 / ((2.0 * np.pi) * INVERSE_H0_MPCH / box) ** 2
    return dict(a=a, phi_crit=phicrit, coupling=coupling, massterm2=massterm2)


def dgp_step_scalars(a, omega, rcH0, rsmooth):
    """beta_DGP (udf:453-455), coupling 1/(3 beta) (udf:593-595), fac0 (udf:762)."""
    H = np.sqrt(omega / (a * a * a) + 1.0 - omega)
    dH = 1.0 / (2.0 * H) * (-3.0 * omega / (a * a * a * a))
    beta = 1.0 + 2.0 * rcH0 * (H + a * dH / 3.0)
    return dict(a=a, coupling=1.0 / (3.0 * beta), dgp_fac0=8.0 / 9.0 * omega * (rcH0 / beta) ** 2, rsmooth=rsmooth)


class ScaleDependentGrowth:
    """Scale-dependent first / second order growth factors: cosmo.c:206-237 (`ode_growth_D` with
    -DSCALEDEPENDENT), 758-1010 (`calculate_scale_dependent_growth_factor`, `growth_X_scaledependent`).

    The ODE system is integrated for `nk` log-spaced wave numbers between the reference's own limits
    (cosmo.c:768-770) with a vectorised fixed-step RK4 in x = ln a; D, dD/dy = D' Q/a,
    ddD/ddy = 1.5 mu Omega a D (and the second-order analogues) are stored on the (x, ln k) grid like the
    reference's 2-D splines and looked up with cubic splines.  `table(field, order, A, AFF)` is what crosses
    the C ABI: the growth factor at every integer m = |d|^2, k = 2 pi sqrt(m) / Box.

    model = "fofr": mu = 1 + 2 beta^2 k^2 / (k^2 + a^2 m^2(a)), beta^2 = 1/6 (udf:618-630, 462-470, 490-498) and the
    Bose-Koyama gamma_2 kernel (udf:848-882); model = "dgp": mu = 1 + 1/(3 beta_DGP) (udf:632-635), gamma_2 of
    udf:884-888; model = "lcdm": mu = 1."""

    def __init__(self, lcdm, box, nmesh, model="fofr", fofr0=1e-5, nfofr=1.0, rcH0=1.0, nk=400, npts=1000, substeps=4):
        self.lcdm, self.box, self.N, self.model = lcdm, float(box), int(nmesh), model
        om = lcdm.omega
        zini = 200.0
        xini, xend = np.log(1.0 / (1.0 + zini)), np.log(2.0)
        kmin = 2.0 * np.pi / box * 0.5                       # cosmo.c:768-769
        kmax = 2.0 * np.pi / box * nmesh * np.sqrt(3.0) * 2.0
        self.logk = np.linspace(np.log(kmin), np.log(kmax), nk)
        k = np.exp(self.logk)
        self.x = np.linspace(xini, xend, npts)

        def mass2(a):
            fac = om / a ** 3 + 4.0 * (1.0 - om)
            fac0 = om + 4.0 * (1.0 - om)
            return fac0 * (fac / fac0) ** (nfofr + 2.0) / ((1.0 + nfofr) * fofr0)

        def beta_dgp(a):
            H = lcdm.hubble(a)
            return 1.0 + 2.0 * rcH0 * (H + a * lcdm.dhubbleda(a) / 3.0)

        def mu_of(a):
            if model == "fofr":
                k2 = (k * INVERSE_H0_MPCH) ** 2
                return 1.0 + 2.0 * (1.0 / 6.0) * k2 / (k2 + a * a * mass2(a))
            if model == "dgp":
                return np.full_like(k, 1.0 + 1.0 / (3.0 * beta_dgp(a)))
            return np.ones_like(k)

        def gamma2_of(a):
            H = lcdm.hubble(a)
            if model == "fofr":
                fac = om / a ** 3 + 4.0 * (1.0 - om)
                fac0 = om + 4.0 * (1.0 - om)

                def pi_fac(kk):
                    return (kk * INVERSE_H0_MPCH / a) ** 2 + fac0 * (fac / fac0) ** (nfofr + 2.0) / (1.0 + nfofr) / fofr0
                g = -(3.0 * (nfofr + 2.0) / (12.0 * (1.0 + nfofr) ** 2)) * (k * INVERSE_H0_MPCH / (a * H)) ** 2 * (om / a ** 3) ** 2 \
                    * fac0 * (fac / fac0) ** (2.0 * nfofr + 3.0) / fofr0 ** 2
                return g / (pi_fac(k) * pi_fac(k / np.sqrt(2.0)) ** 2)
            if model == "dgp":
                return np.full_like(k, -1.0 / 6.0 / beta_dgp(a) ** 3 * (rcH0 / H) ** 2 * (om / a ** 3) ** 2)
            return np.zeros_like(k)

        self._mu_of = mu_of

        def rhs(xx, y):
            a = np.exp(xx)
            H, dH = lcdm.hubble(a), lcdm.dhubbleda(a)
            mu = mu_of(a)
            beta = 1.5 * om * mu / (a ** 3 * H * H)
            alpha = 2.0 + a * dH / H
            modfac = 1.0 + 2.0 * (a * a * H) ** 2 * gamma2_of(a) / (1.5 * om * a * mu)
            return np.stack([y[1], -alpha * y[1] + beta * y[0], y[3], -alpha * y[3] + beta * (y[2] - modfac * y[0] * y[0])])

        y = np.stack([np.ones(nk), np.ones(nk), np.full(nk, -3.0 / 7.0), np.full(nk, -6.0 / 7.0)])
        store = np.empty((npts, 4, nk))
        store[0] = y
        for i in range(1, npts):
            h = (self.x[i] - self.x[i - 1]) / substeps
            xx = self.x[i - 1]
            for _ in range(substeps):
                k1 = rhs(xx, y)
                k2_ = rhs(xx + 0.5 * h, y + 0.5 * h * k1)
                k3 = rhs(xx + 0.5 * h, y + 0.5 * h * k2_)
                k4 = rhs(xx + h, y + h * k3)
                y = y + h / 6.0 * (k1 + 2 * k2_ + 2 * k3 + k4)
                xx += h
            store[i] = y
        a = np.exp(self.x)[:, None]
        Q = lcdm.qfactor(a)
        mu = np.stack([mu_of(aa) for aa in np.exp(self.x)])
        D, q, D2, q2 = store[:, 0], store[:, 1], store[:, 2], store[:, 3]
        self._tab = {("D", 1): D, ("dD", 1): q * Q / a, ("ddD", 1): 1.5 * mu * om * a * D,
                     ("D", 2): D2, ("dD", 2): q2 * Q / a, ("ddD", 2): 1.5 * mu * om * a * (D2 - D * D)}
        self._spl = {kk: CubicSpline(self.x, v, axis=0) for kk, v in self._tab.items()}
        self._norm = {1: self._spl[("D", 1)](0.0), 2: self._spl[("D", 2)](0.0)}     # value at a = 1 per k (cosmo.c:987-1010)
        h2 = (nmesh // 2) ** 2
        m = np.arange(3 * h2 + 1, dtype=np.float64)
        with np.errstate(divide="ignore"):
            self._logk_m = np.log(2.0 * np.pi / box * np.sqrt(m))
        self._logk_m[0] = self.logk[0]

    def _column(self, name, order, a):
        return self._spl[(name, order)](np.log(a)) / self._norm[order]

    def of_k2(self, name, order, a):
        """growth_<name>_scaledependent(k(m), a) for every integer m = |d|^2 (entry 0 is never used)."""
        col = self._column(name, order, a)
        out = CubicSpline(self.logk, col)(np.clip(self._logk_m, self.logk[0], self.logk[-1]))
        out[0] = 0.0
        return out

    def table(self, fieldtype, order, A, AFF=None):
        """The growth factor from_cdisp_store_to_ZA applies (2LPT.c:1611-1614), without its normfactor."""
        if fieldtype == 0:
            return self.of_k2("D", order, A)
        if fieldtype == 1:
            return self.of_k2("dD", order, A)
        if fieldtype == 2:
            return self.of_k2("ddD", order, A)
        return self.of_k2("D", order, AFF) - self.of_k2("D", order, A)

    def pofk_ratio_by_k2(self):
        """mg_pofk_ratio(k, 1) = (D(k, 1) / D_LCDM(1))^2 with both unnormalised (cosmo.c:708-714)."""
        col = self._norm[1] / self.lcdm._n1
        out = CubicSpline(self.logk, col)(np.clip(self._logk_m, self.logk[0], self.logk[-1])) ** 2
        out[0] = 0.0
        return out
