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
