"""Synthetic worlds of BASELINE.json's large-N configs, generated identically for the CUDA path
and the CPU oracle (numpy, FP64, fixed seeds).

* plummer(n, seed)            -- Aarseth-Henon-Wielen sampling of a Plummer sphere, equal masses
                                 m = 2e30 kg, scale radius a = 1.5e14 m, radii truncated at 10 a,
                                 recentred to zero centre-of-mass position and momentum.
* galaxy_collision(n, seed)   -- two Plummer spheres of n/2 bodies on a bound, off-axis encounter.
* default_dt(n)               -- the integer tick length (World::m_simulation_seconds_per_tick is
                                 an int, reference src/World.hpp:78) ~ 1e-3 dynamical times.
"""
import numpy as np

GRAVITY = 6.67408e-11
BODY_MASS = 2e30
SCALE_RADIUS = 1.5e14
SEED = 20240420


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=int(seed)))


def _isotropic(rng, n):
    cos_t = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    sin_t = np.sqrt(1.0 - cos_t * cos_t)
    return sin_t * np.cos(phi), sin_t * np.sin(phi), cos_t


def plummer(n, seed=SEED, mass=BODY_MASS, a=SCALE_RADIUS):
    """Returns dict(x, y, z, vx, vy, vz, gf) of float64 arrays (SI units)."""
    rng = _rng(seed)
    # radii: r = a / sqrt(X^(-2/3) - 1), resampled beyond 10 a
    r = np.empty(n)
    todo = np.arange(n)
    while len(todo):
        xm = rng.uniform(0.0, 1.0, len(todo))
        xm = np.clip(xm, 1e-12, 1.0 - 1e-12)
        rr = 1.0 / np.sqrt(xm ** (-2.0 / 3.0) - 1.0)
        ok = rr <= 10.0
        r[todo[ok]] = rr[ok]
        todo = todo[~ok]
    ux, uy, uz = _isotropic(rng, n)
    x, y, z = r * ux, r * uy, r * uz
    # speeds: q = v / v_esc from g(q) = q^2 (1 - q^2)^(7/2) by von Neumann rejection
    q = np.empty(n)
    todo = np.arange(n)
    while len(todo):
        x4 = rng.uniform(0.0, 1.0, len(todo))
        x5 = rng.uniform(0.0, 1.0, len(todo))
        ok = 0.1 * x5 < x4 * x4 * (1.0 - x4 * x4) ** 3.5
        q[todo[ok]] = x4[ok]
        todo = todo[~ok]
    v = q * np.sqrt(2.0) * (1.0 + r * r) ** -0.25
    wx, wy, wz = _isotropic(rng, n)
    vx, vy, vz = v * wx, v * wy, v * wz
    # N-body units (G = M = a = 1) -> SI
    gm = GRAVITY * mass * n
    vscale = np.sqrt(gm / a)
    out = dict(x=x * a, y=y * a, z=z * a, vx=vx * vscale, vy=vy * vscale, vz=vz * vscale)
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        out[k] = out[k] - out[k].mean()  # equal masses: mean = centre of mass
    out["gf"] = np.full(n, mass * GRAVITY)
    return out


def galaxy_collision(n, seed=SEED, mass=BODY_MASS, a=SCALE_RADIUS):
    """Two Plummer spheres of n/2 bodies each, centres at (-+4a, -+a, 0), bulk velocities
    +-0.5 sqrt(G M_tot / 8a) along x (bound, off-axis)."""
    h = n // 2
    p1 = plummer(h, seed + 1, mass, a)
    p2 = plummer(n - h, seed + 2, mass, a)
    vb = 0.5 * np.sqrt(GRAVITY * mass * n / (8.0 * a))
    out = {}
    for k in ("x", "y", "z", "vx", "vy", "vz", "gf"):
        out[k] = np.concatenate([p1[k], p2[k]])
    out["x"][:h] -= 4.0 * a
    out["x"][h:] += 4.0 * a
    out["y"][:h] -= a
    out["y"][h:] += a
    out["vx"][:h] += vb
    out["vx"][h:] -= vb
    return out


def default_dt(n, a=SCALE_RADIUS, mass=BODY_MASS):
    """Integer seconds ~ 1e-3 sqrt(a^3 / G M)."""
    t_dyn = np.sqrt(a ** 3 / (GRAVITY * mass * n))
    dt = 1e-3 * t_dyn
    mag = 10.0 ** np.floor(np.log10(dt))
    return int(max(1, round(dt / mag) * mag))
