"""Synthetic pressure-step inputs for the BASELINE.json configurations.

Pure numpy, host side: grid/BC description, seeded velocity fields made BC-consistent so that the
Poisson RHS is solvable, and the ghost-cell updates a Fortran main does around the path
(`boundp`, src/bound.f90:146-225 + set_bc :227-646, restated for the homogeneous pressure BCs that
FluTAS allows, sanity.f90:238-244).
"""
import numpy as np

from .initsolver import SolverSetup

# id -> (ng, pressure BCs x/y/z, lengths, rho0, seed offset)
CONFIGS = {
    "C1": dict(ng=(64, 64, 64), cbc=("PP", "PP", "PP"), l=(2 * np.pi,) * 3, rho0=1.0, idx=1,
               desc="single-phase tri-periodic 64^3 parity case"),
    "C2": dict(ng=(512, 512, 512), cbc=("PP", "PP", "PP"), l=(2 * np.pi,) * 3, rho0=1.0, idx=2,
               desc="HIT 512^3 tri-periodic FP64 (DFT x,y + periodic z)"),
    "C3": dict(ng=(1024, 512, 512), cbc=("PP", "PP", "NN"), l=(6.0, 3.0, 1.0), rho0=1.0, idx=3,
               desc="turbulent channel 1024x512x512, Neumann walls in z"),
    "C4": dict(ng=(512, 512, 512), cbc=("PP", "PP", "PP"), l=(2 * np.pi,) * 3, rho0=0.1, idx=4,
               desc="two-phase droplet-laden HIT 512^3, constant-coefficient split (rho0=min(rho1,rho2))"),
    "C5": dict(ng=(2048, 2048, 1024), cbc=("NN", "NN", "NN"), l=(2.0, 2.0, 1.0), rho0=1.0, idx=5,
               desc="Rayleigh-Benard 2048x2048x1024, DCT x,y + walls z"),
    "C5w1": dict(ng=(1024, 1024, 512), cbc=("NN", "NN", "NN"), l=(2.0, 2.0, 1.0), rho0=1.0, idx=5,
                 desc="Rayleigh-Benard weak-scaling unit 1024x1024x512 per GPU"),
    # one eighth (a z-slab of 128 levels) of C5: the x / y transform kernels at N = 2048 on ONE GPU (kernel timing only;
    # its z stage sees nz = 128 and says nothing about C5's)
    "C5xy": dict(ng=(2048, 2048, 128), cbc=("NN", "NN", "NN"), l=(2.0, 2.0, 0.125), rho0=1.0, idx=5,
                 desc="x/y stages of Rayleigh-Benard 2048x2048x1024 on one z-slab of 128 levels"),
    "NS": dict(ng=(1024, 1024, 1024), cbc=("PP", "PP", "NN"), l=(2 * np.pi, 2 * np.pi, 1.0), rho0=1.0, idx=6,
               desc="north-star 1024^3 channel"),
}
SEED0 = 20261017


class Case:
    """One pressure-step problem on a single rank (x-pencil == whole domain)."""

    def __init__(self, ng, cbc, lengths, rho0=1.0, dt=1.0e-3, gr=0.0, nh_u=1, seed=SEED0, name="custom"):
        self.name = name
        self.ng = tuple(int(x) for x in ng)
        self.cbc = tuple(cbc)
        self.lengths = tuple(float(x) for x in lengths)
        self.rho0 = float(rho0)
        self.dt = float(dt)
        self.dti = 1.0 / self.dt
        self.nh_u = int(nh_u)
        self.nh_d = max(1, self.nh_u)
        self.nh_p = 1
        self.seed = int(seed)
        self.setup = SolverSetup(self.ng, self.lengths, self.cbc, gr=gr, nh_d=self.nh_d)

    @classmethod
    def from_config(cls, cid, ng=None, **kw):
        c = CONFIGS[cid]
        return cls(ng or c["ng"], c["cbc"], c["l"], rho0=c["rho0"], seed=SEED0 + c["idx"], name=cid, **kw)

    # ---- fields -------------------------------------------------------------------------------
    def velocity(self):
        """u,v,w ~ U(-1,1) from PCG64(seed), halos made consistent with the pressure BCs."""
        n1, n2, n3 = self.ng
        h = self.nh_u
        rng = np.random.Generator(np.random.PCG64(self.seed))
        shape = (n1 + 2 * h, n2 + 2 * h, n3 + 2 * h)
        out = []
        for _ in range(3):
            f = np.zeros(shape, order="F")
            f[...] = rng.uniform(-1.0, 1.0, size=shape[::-1]).T
            out.append(f)
        self.refresh_velocity_halos(*out)
        return out

    def refresh_velocity_halos(self, u, v, w):
        """Minimal stand-in for `bounduvw` (src/bound.f90:17-144): periodic wrap, and zero wall-normal
        face velocity at walls where the pressure BC is Neumann (no-penetration)."""
        h = self.nh_u
        for d, n in enumerate(self.ng):
            bc = self.cbc[d]
            fld = (u, v, w)[d]
            for f in (u, v, w):
                if bc == "PP":
                    lo = [slice(None)] * 3
                    hi = [slice(None)] * 3
                    src_lo = [slice(None)] * 3
                    src_hi = [slice(None)] * 3
                    lo[d] = slice(0, h)                  # indices 1-h..0
                    src_lo[d] = slice(n, n + h)          # indices n-h+1..n
                    hi[d] = slice(n + h, n + 2 * h)      # indices n+1..n+h
                    src_hi[d] = slice(h, 2 * h)          # indices 1..h
                    f[tuple(lo)] = f[tuple(src_lo)]
                    f[tuple(hi)] = f[tuple(src_hi)]
            for ib in (0, 1):
                if bc[ib] == "N":                        # wall: normal face velocity = 0
                    sl = [slice(None)] * 3
                    sl[d] = (h - 1) if ib == 0 else (n + h - 1)   # face index 0 / n
                    fld[tuple(sl)] = 0.0

    def new_p(self):
        n1, n2, n3 = self.ng
        return np.zeros((n1 + 2, n2 + 2, n3 + 2), order="F")

    def boundp(self, p):
        """Ghost cells of p for homogeneous pressure BCs (src/bound.f90:247-268,320-420): P wrap,
        N: ghost = inner, D (cell-centred): ghost = -inner."""
        for d, n in enumerate(self.ng):
            bc = self.cbc[d]
            lo = [slice(None)] * 3
            hi = [slice(None)] * 3
            first = [slice(None)] * 3
            last = [slice(None)] * 3
            lo[d], hi[d], first[d], last[d] = 0, n + 1, 1, n
            lo, hi, first, last = tuple(lo), tuple(hi), tuple(first), tuple(last)
            if bc == "PP":
                p[lo] = p[last]
                p[hi] = p[first]
            else:
                p[lo] = p[first] if bc[0] == "N" else -p[first]
                p[hi] = p[last] if bc[1] == "N" else -p[last]
        return p

    def correct_dirichlet_faces(self, p, u, v, w):
        """At a Dirichlet-pressure boundary the face velocity at index 0 lies outside `correc`'s loop;
        a consistent projection corrects it with the same formula (ghost p from boundp)."""
        h = self.nh_u
        s = self.setup
        for d in range(3):
            if self.cbc[d][0] != "D":
                continue
            f = (u, v, w)[d]
            fs = [slice(h, -h)] * 3
            ps0 = [slice(1, -1)] * 3
            ps1 = [slice(1, -1)] * 3
            fs[d], ps0[d], ps1[d] = h - 1, 0, 1
            fac = self.dt * (s.dli[d] if d < 2 else s.dzci[0 + s.nh_d - 1])
            f[tuple(fs)] -= fac * (p[tuple(ps1)] - p[tuple(ps0)]) / self.rho0

    # ---- discrete operator (for residual checks) -----------------------------------------------
    def laplacian(self, p):
        """7-point Laplacian of the interior of p (ghosts must be valid): what `solver` inverts."""
        s = self.setup
        n1, n2, n3 = self.ng
        o = s.nh_d - 1
        c = p[1:-1, 1:-1, 1:-1]
        lap = (p[2:, 1:-1, 1:-1] - 2 * c + p[:-2, 1:-1, 1:-1]) * s.dli[0] ** 2
        lap = lap + (p[1:-1, 2:, 1:-1] - 2 * c + p[1:-1, :-2, 1:-1]) * s.dli[1] ** 2
        k = np.arange(1, n3 + 1)
        az = (s.dzfi[k + o] * s.dzci[k - 1 + o])[None, None, :]
        cz = (s.dzfi[k + o] * s.dzci[k + o])[None, None, :]
        lap = lap + cz * (p[1:-1, 1:-1, 2:] - c) - az * (c - p[1:-1, 1:-1, :-2])
        return lap

    @property
    def singular(self):
        return all(b in ("PP", "NN") for b in self.cbc)


def rel_err_gauge_fixed(p, pref, singular=True):
    """max|dp|/max|p| after removing the mean of each field when the problem is singular
    (SURVEY.md section 7 hard part 1); also returns the raw figure."""
    a = p[1:-1, 1:-1, 1:-1]
    b = pref[1:-1, 1:-1, 1:-1]
    raw = np.max(np.abs(a - b)) / np.max(np.abs(b))
    if singular:
        a0, b0 = a - a.mean(), b - b.mean()
        return np.max(np.abs(a0 - b0)) / np.max(np.abs(b0)), raw
    return raw, raw
