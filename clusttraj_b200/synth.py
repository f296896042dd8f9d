"""Synthetic trajectories of the BASELINE.json shapes.

No dataset can be downloaded, so the bench and the parity tests use seeded
synthetic frames: a base structure, a handful of conformational modes, thermal
noise, then a random proper rotation and a translation per frame so centring
and superposition are really exercised.  Coordinates are rounded to 5
decimals like an XYZ file.  Pure numpy; no CUDA.
"""

import numpy as np


def _random_rotations(rng, m):
    """m proper rotation matrices from uniformly random unit quaternions."""
    q = rng.normal(size=(m, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((m, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _apply_rigid(rng, frames, translate=20.0):
    m = len(frames)
    R = _random_rotations(rng, m)
    out = np.einsum("mnk,mjk->mnj", frames, R)
    out += rng.uniform(-translate, translate, size=(m, 1, 3))
    return out


def make_elements(n_atoms, h_fraction=0.5, seed=0):
    """Frame-invariant atomic numbers: heavy atoms (C/N/O) interleaved with H."""
    rng = np.random.default_rng(seed + 7)
    z = np.empty(n_atoms, dtype=np.int32)
    n_h = int(round(n_atoms * h_fraction))
    is_h = np.zeros(n_atoms, dtype=bool)
    if n_h:
        is_h[np.round(np.linspace(0, n_atoms - 1, n_h)).astype(int)] = True
        # linspace rounding can collide; top up deterministically
        k = 0
        while is_h.sum() < n_h:
            if not is_h[k]:
                is_h[k] = True
            k += 1
    z[is_h] = 1
    z[~is_h] = rng.choice(np.array([6, 7, 8], dtype=np.int32), size=int((~is_h).sum()), p=[0.6, 0.2, 0.2])
    return z


def make_trajectory(n_frames, n_atoms, seed=20250601, n_modes=8, noise=0.3, mode_spread=1.0,
                    h_fraction=0.5, n_duplicates=2, mirrored=True, decimals=5):
    """Plain / ``-n`` workloads (BASELINE configs 2, 3, 5).

    Returns (atomic_numbers int32[N], coords float64[M,N,3]).
    Frames 1..n_duplicates are exact rotated+translated copies of frame 0 up to
    the 5-decimal rounding (RMSD ~ 1e-5 edge); with ``mirrored`` the last frame
    is the mirror image of frame 0 (det C < 0 edge).
    """
    rng = np.random.default_rng(seed)
    scale = 1.5 * n_atoms ** (1.0 / 3.0)
    base = rng.normal(0.0, scale, size=(n_atoms, 3))
    centres = base[None] + rng.normal(0.0, mode_spread, size=(n_modes, n_atoms, 3))
    which = rng.integers(0, n_modes, size=n_frames)
    frames = centres[which]
    frames = frames + rng.normal(0.0, noise, size=frames.shape)
    for d in range(1, min(n_duplicates, n_frames - 1) + 1):
        frames[d] = frames[0]
    if mirrored and n_frames > n_duplicates + 2:
        frames[-1] = frames[0] * np.array([1.0, 1.0, -1.0])
    frames = _apply_rigid(rng, frames)
    if decimals is not None:
        frames = np.round(frames, decimals)
    return make_elements(n_atoms, h_fraction, seed), np.ascontiguousarray(frames)


def make_solute_solvent(n_frames, n_solute=60, n_solvent_mol=100, seed=20250604, n_modes=4,
                        noise=0.15, decimals=5, box=14.0):
    """Solute + rigid three-atom (O,H,H) solvent molecules (BASELINE config 4).

    Every frame scatters the solvent molecules over the same set of sites (per
    mode) with thermal noise, shuffles which molecule sits on which site and
    randomly swaps the two hydrogens, so that the Hungarian reorder has real
    work to do and a unique (non-degenerate) answer.
    Returns (atomic_numbers int32[N], coords float64[M,N,3]) with the solute
    first, N = n_solute + 3*n_solvent_mol.
    """
    rng = np.random.default_rng(seed)
    z_sol = rng.choice(np.array([6, 7, 8, 1], dtype=np.int32), size=n_solute, p=[0.4, 0.1, 0.15, 0.35])
    z_sol[:3] = (6, 7, 8)
    sol_base = rng.normal(0.0, 1.5 * n_solute ** (1.0 / 3.0), size=(n_solute, 3))
    sol_modes = sol_base[None] + rng.normal(0.0, 0.6, size=(n_modes, n_solute, 3))

    # solvent sites on a jittered grid well separated (> 2.5 A) from each other
    g = int(np.ceil(n_solvent_mol ** (1.0 / 3.0))) + 1
    axis = np.linspace(-box, box, g)
    grid = np.stack(np.meshgrid(axis, axis, axis, indexing="ij"), -1).reshape(-1, 3)
    grid = grid[rng.permutation(len(grid))[:n_solvent_mol]]
    sites = grid[None] + rng.normal(0.0, 0.5, size=(n_modes, n_solvent_mol, 3))
    water = np.array([[0.0, 0.0, 0.1173], [0.0, 0.7572, -0.4692], [0.0, -0.7572, -0.4692]])
    orient = _random_rotations(rng, n_modes * n_solvent_mol).reshape(n_modes, n_solvent_mol, 3, 3)

    which = rng.integers(0, n_modes, size=n_frames)
    solute = sol_modes[which] + rng.normal(0.0, noise, size=(n_frames, n_solute, 3))
    centres = sites[which] + rng.normal(0.0, noise, size=(n_frames, n_solvent_mol, 3))
    mol = np.einsum("ak,mskj->msaj", water, np.swapaxes(orient[which], -1, -2))
    mol = mol + centres[:, :, None, :] + rng.normal(0.0, 0.05, size=(n_frames, n_solvent_mol, 3, 3))
    # shuffle molecule labels and swap hydrogens per frame
    for m in range(n_frames):
        mol[m] = mol[m, rng.permutation(n_solvent_mol)]
    swap = rng.random((n_frames, n_solvent_mol)) < 0.5
    h1 = mol[:, :, 1, :].copy()
    mol[:, :, 1, :] = np.where(swap[..., None], mol[:, :, 2, :], mol[:, :, 1, :])
    mol[:, :, 2, :] = np.where(swap[..., None], h1, mol[:, :, 2, :])
    frames = np.concatenate([solute, mol.reshape(n_frames, 3 * n_solvent_mol, 3)], axis=1)
    frames = _apply_rigid(rng, frames)
    if decimals is not None:
        frames = np.round(frames, decimals)
    z = np.concatenate([z_sol, np.tile(np.array([8, 1, 1], dtype=np.int32), n_solvent_mol)])
    return z.astype(np.int32), np.ascontiguousarray(frames)
