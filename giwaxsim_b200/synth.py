"""Synthetic inputs for benchmarks and tests: seeded random-atom slabs and a
fixed f'/f'' table (xraydb is not vendored; the hot path takes f-values as an
input, so any fixed table exercises it identically)."""
import numpy as np

# approximate Chantler f', f'' near 12.7 keV -- placeholders, see DESIGN.md
F1F2_TABLE = {
    "H": (0.0, 0.0), "C": (0.0049, 0.0023), "N": (0.0090, 0.0047), "O": (0.0160, 0.0090),
    "F": (0.0240, 0.0140), "Si": (0.1100, 0.1000), "P": (0.1400, 0.1400), "S": (0.1700, 0.2500),
    "Cl": (0.2000, 0.2400),
}


def fixed_f1f2(element, energy=None):
    return F1F2_TABLE[element]


def random_slab(n_atoms, box, seed=20240829, elements=("C", "H", "S", "O", "F"),
                fractions=(0.62, 0.30, 0.04, 0.02, 0.02)):
    """Uniform random atoms in an orthorhombic box (Angstrom); returns
    (coords float64 [A,3], elements '<U2' [A])."""
    rng = np.random.default_rng(seed)
    coords = rng.random((int(n_atoms), 3)) * np.asarray(box, dtype=np.float64)
    el = rng.choice(np.asarray(elements), size=int(n_atoms), p=np.asarray(fractions))
    return coords, el


def pow2_q_voxel(r_voxel_size, grid_size):
    """q_voxel_size for which the reference's ceil(2 pi/(q r)) is exactly grid_size."""
    return 2 * np.pi / (r_voxel_size * (grid_size - 0.5))


def config5(scale=1.0):
    """BASELINE.json configs[4]: ~10 M random atoms, 4096^2 grid, q_voxel 0.01,
    max_q 2 -> q_num 569; 2048^2 detector over 360 psi.  `scale` < 1 shrinks the
    atom count only (same grid), for bounded CPU-baseline samples."""
    q = 0.01
    N = 4096
    r = 2 * np.pi / (q * (N - 0.5))
    return dict(n_atoms=int(10_000_000 * scale), box=(560.0, 250.0, 560.0), r_voxel_size=r,
                q_voxel_size=q, max_q=2.0, grid_size=N, fill_bkg=True, smooth=25, energy=12700.0,
                num_pixels=2048, psis=np.linspace(0, 89.75, 360), phis=np.array([0.0]),
                thetas=np.array([0.0]), angle_init_vals=(90.0, 90.0, 90.0),
                angle_init_axs=("psi", "phi", "psi"), n_phi=1800)
