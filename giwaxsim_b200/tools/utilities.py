"""Helpers the hot-path entry points share: device "shared arrays" standing in
for the reference's POSIX shared-memory accumulators, the scattering-factor
lookup hook, and the axis-angle matrix.

Reference: tools/utilities.py:222-245 (rotation_matrix), :342-366
(get_element_f1_f2_dict), :378-398 (create_shared_array).
"""
import uuid

import numpy as np
import torch

ATOMIC_NUMBER = {
    'H': 1, 'He': 2, 'Li': 3, 'Be': 4, 'B': 5, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'Ne': 10, 'Na': 11,
    'Mg': 12, 'Al': 13, 'Si': 14, 'P': 15, 'S': 16, 'Cl': 17, 'Ar': 18, 'K': 19, 'Ca': 20, 'Sc': 21,
    'Ti': 22, 'V': 23, 'Cr': 24, 'Mn': 25, 'Fe': 26, 'Co': 27, 'Ni': 28, 'Cu': 29, 'Zn': 30, 'Ga': 31,
    'Ge': 32, 'As': 33, 'Se': 34, 'Br': 35, 'Kr': 36, 'Rb': 37, 'Sr': 38, 'Y': 39, 'Zr': 40, 'Nb': 41,
    'Mo': 42, 'Tc': 43, 'Ru': 44, 'Rh': 45, 'Pd': 46, 'Ag': 47, 'Cd': 48, 'In': 49, 'Sn': 50, 'Sb': 51,
    'Te': 52, 'I': 53, 'Xe': 54, 'Cs': 55, 'Ba': 56, 'La': 57, 'Ce': 58, 'Pr': 59, 'Nd': 60, 'Pm': 61,
    'Sm': 62, 'Eu': 63, 'Gd': 64, 'Tb': 65, 'Dy': 66, 'Ho': 67, 'Er': 68, 'Tm': 69, 'Yb': 70, 'Lu': 71,
    'Hf': 72, 'Ta': 73, 'W': 74, 'Re': 75, 'Os': 76, 'Ir': 77, 'Pt': 78, 'Au': 79, 'Hg': 80, 'Tl': 81,
    'Pb': 82, 'Bi': 83, 'Po': 84, 'At': 85, 'Rn': 86, 'Fr': 87, 'Ra': 88, 'Ac': 89, 'Th': 90, 'Pa': 91,
    'U': 92}

# Cromer-Mann a1,b1,...,a4,b4,c (International Tables C, table 6.1.1.4) of the
# neutral atoms that occur in organic / hybrid films; the fitting path itself only
# ever uses carbon (comparison.py:785), the two-step driver the most common element
# of the input file (old_modules/voxelgridmaker.py:66-68).  Extend with
# register_cromer_mann for anything else (the reference's own table has no neutral
# 'Si' entry, for instance: ptable_dict.py).
CROMER_MANN = {
    'H': (0.489918, 20.6593, 0.262003, 7.74039, 0.196767, 49.5519, 0.049879, 2.20159, 0.001305),
    'B': (2.0545, 23.2185, 1.3326, 1.021, 1.0979, 60.3498, 0.7068, 0.1403, -0.1932),
    'C': (2.31, 20.8439, 1.02, 10.2075, 1.5886, 0.5687, 0.865, 51.6512, 0.2156),
    'N': (12.2126, 0.0057, 3.1322, 9.8933, 2.0125, 28.9975, 1.1663, 0.5826, -11.529),
    'O': (3.0485, 13.2771, 2.2868, 5.7011, 1.5463, 0.3239, 0.867, 32.9089, 0.2508),
    'F': (3.5392, 10.2825, 2.6412, 4.2944, 1.517, 0.2615, 1.0243, 26.1476, 0.2776),
    'Na': (4.7626, 3.285, 3.1736, 8.8422, 1.2674, 0.3136, 1.1128, 129.424, 0.676),
    'Mg': (5.4204, 2.8275, 2.1735, 79.2611, 1.2269, 0.3808, 2.3073, 7.1937, 0.8584),
    'Al': (6.4202, 3.0387, 1.9002, 0.7426, 1.5936, 31.5472, 1.9646, 85.0886, 1.1151),
    'P': (6.4345, 1.9067, 4.1791, 27.157, 1.78, 0.526, 1.4908, 68.1645, 1.1149),
    'S': (6.9053, 1.4679, 5.2034, 22.2151, 1.4379, 0.2536, 1.5863, 56.172, 0.8669),
    'Cl': (11.4604, 0.0104, 7.1964, 1.1662, 6.2556, 18.5194, 1.6455, 47.7784, -9.5574),
    'K': (8.2186, 12.7949, 7.4398, 0.7748, 1.0519, 213.187, 0.8659, 41.6841, 1.4228),
    'Ca': (8.6266, 10.4421, 7.3873, 0.6599, 1.5899, 85.7484, 1.0211, 178.437, 1.3751),
    'Ti': (9.7595, 7.8508, 7.3558, 0.5, 1.6991, 35.6338, 1.9021, 116.105, 1.2807),
    'Fe': (11.7695, 4.7611, 7.3573, 0.3072, 3.5222, 15.3535, 2.3045, 76.8805, 1.0369),
    'Cu': (13.338, 3.5828, 7.1676, 0.247, 5.6158, 11.3966, 1.6735, 64.8126, 1.191),
    'Zn': (14.0743, 3.2655, 7.0318, 0.2333, 5.1652, 10.3163, 2.41, 58.7097, 1.3041),
    'Se': (17.0006, 2.4098, 5.8196, 0.2726, 3.9731, 15.2372, 4.3543, 43.8163, 2.8409),
    'Br': (17.1789, 2.1723, 5.2358, 16.5796, 5.6377, 0.2609, 3.9851, 41.4328, 2.9557),
    'I': (20.1472, 4.347, 18.9949, 0.3814, 7.5138, 27.766, 2.2735, 66.8776, 4.0712),
    'Au': (16.8819, 0.4611, 18.5913, 8.6216, 25.5582, 1.4826, 5.86, 36.3956, 12.0658),
    'Pb': (31.0617, 0.6902, 13.0637, 2.3576, 18.442, 8.618, 5.9696, 47.2579, 13.4118),
}


def register_cromer_mann(element, coefficients):
    """Add / replace the nine Cromer-Mann coefficients (a1,b1,..,a4,b4,c) of an element."""
    coefficients = tuple(float(v) for v in coefficients)
    if len(coefficients) != 9:
        raise ValueError("nine coefficients expected: a1, b1, a2, b2, a3, b3, a4, b4, c")
    CROMER_MANN[str(element)] = coefficients


def get_element_f0_dict(q_val, elements):
    """{element: f0} of the distinct elements at q_val (utilities.py:319-340).  The
    reference's exponent is -b*q/(16 pi^2) - q, not q^2 - and this mirrors it."""
    out = {}
    for element in set(elements):
        aff = CROMER_MANN[element]
        out[element] = (aff[0] * np.exp(-aff[1] * (q_val) / (16 * np.pi ** 2)) +
                        aff[2] * np.exp(-aff[3] * (q_val) / (16 * np.pi ** 2)) +
                        aff[4] * np.exp(-aff[5] * (q_val) / (16 * np.pi ** 2)) +
                        aff[6] * np.exp(-aff[7] * (q_val) / (16 * np.pi ** 2)) +
                        aff[8])
    return out

_f1f2_provider = None


def set_f1f2_provider(fn):
    """Install fn(element, energy_eV) -> (f', f'') used instead of xraydb."""
    global _f1f2_provider
    _f1f2_provider = fn


def get_element_f1_f2_dict(energy, elements):
    """{element: f' + i f''} for the distinct elements (utilities.py:342-366).
    Uses xraydb's Chantler tables like the reference unless a provider was
    installed with set_f1f2_provider (the test/bench harness does, because
    xraydb is not vendored)."""
    out = {}
    provider = _f1f2_provider
    if provider is None:
        try:
            import xraydb
        except ImportError as e:
            raise ImportError("xraydb is required for the f'/f'' lookup (or call "
                              "giwaxsim_b200.tools.utilities.set_f1f2_provider)") from e

        def provider(el, en):
            return xraydb.f1_chantler(el, en), xraydb.f2_chantler(el, en)
    for element in set(elements):
        try:
            f1, f2 = provider(element, energy)
            out[element] = f1 + 1j * f2
        except KeyError:
            print(f"Data for element '{element}' at energy {energy} eV not found.")
    return out


def rotation_matrix(u, theta):
    """Axis-angle (Rodrigues) matrix, operations associated as in
    utilities.py:222-245 so the entries are bit-identical."""
    ux, uy, uz = u[0], u[1], u[2]
    c, s = np.cos(theta), np.sin(theta)
    k = 1 - c
    return np.array([[c + ux ** 2 * k, ux * uy * k - uz * s, ux * uz * k + uy * s],
                     [uy * ux * k + uz * s, c + uy ** 2 * k, uy * uz * k - ux * s],
                     [uz * ux * k - uy * s, uz * uy * k + ux * s, c + uz ** 2 * k]])


# ---------------------------------------------------------------------------
# device accumulators addressed by name
# ---------------------------------------------------------------------------
_registry = {}


class DeviceSharedArray:
    """Stand-in for multiprocessing.shared_memory.SharedMemory on the GPU.

    The reference creates float64 POSIX shared-memory blocks and hands their
    *names* to the workers (comparison.py:747-750, :835-841).  Here the name
    keys a device tensor owned by this object; its dtype is chosen by the first
    worker that uses it (fp32 intensity sums, int32 counts, fp64 detector
    image).  `.buf` returns a float64 host snapshot so the reference idiom
    `np.ndarray(shape, dtype=np.float64, buffer=shm.buf)` keeps working for
    reading results.
    """

    def __init__(self, shape, name=None):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.name = name or uuid.uuid4().hex[:29]
        self.size = int(np.prod(self.shape)) * 8
        self.tensor = None
        self.aux = {}
        _registry[self.name] = self

    def device_tensor(self, dtype, device):
        if self.tensor is None:
            self.tensor = torch.zeros(int(np.prod(self.shape)), dtype=dtype, device=device)
        elif self.tensor.dtype != dtype:
            raise TypeError("shared array %s already holds %s, asked for %s"
                            % (self.name, self.tensor.dtype, dtype))
        return self.tensor

    def to_numpy(self):
        if self.tensor is None:
            return np.zeros(self.shape, dtype=np.float64)
        return self.tensor.detach().cpu().to(torch.float64).numpy().reshape(self.shape)

    @property
    def buf(self):
        return memoryview(np.ascontiguousarray(self.to_numpy())).cast("B")

    def close(self):
        pass

    def unlink(self):
        _registry.pop(self.name, None)
        self.tensor = None
        self.aux = {}


def create_shared_array(shape, name=None):
    """utilities.py:378-398: zero-filled accumulator; here on the device."""
    return DeviceSharedArray(shape, None)


def lookup_shared_array(name):
    try:
        return _registry[name]
    except KeyError:
        raise FileNotFoundError("no device shared array named %r" % (name,)) from None


# ---------------------------------------------------------------------------
# structure files and cell vectors (inputs of slabmaker_fitting)
# ---------------------------------------------------------------------------
def _symbol(token):
    """Element symbol of an .xyz first column: digits removed (utilities.py:52-55 use)."""
    return "".join(ch for ch in token if not ch.isdigit())


def load_xyz(xyz_path):
    """(coords [n,3] float64, elements [n]) of an XYZ file: two header lines, then
    `symbol x y z ...` per atom; malformed lines are skipped (utilities.py:97-132)."""
    coords, symbols = [], []
    with open(xyz_path, "r") as fh:
        body = fh.readlines()[2:]
    for line in body:
        parts = line.split()
        if len(parts) < 4:
            continue
        try:
            xyz = [float(parts[1]), float(parts[2]), float(parts[3])]
        except ValueError:
            continue
        symbols.append(_symbol(parts[0]))
        coords.append(xyz)
    return np.array(coords), np.array(symbols)


def load_pdb(pdb_path):
    """(coords, elements) of the ATOM / HETATM records of a PDB file: fixed columns
    31-54 for x, y, z and 77-78 for the element (utilities.py:134-161)."""
    coords, symbols = [], []
    with open(pdb_path, "r") as fh:
        for line in fh:
            if line.startswith(("ATOM", "HETATM")):
                symbols.append(line[76:78].strip())
                coords.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
    return np.array(coords), np.array(symbols)


def load_structure(input_path):
    """load_xyz / load_pdb by the last three characters of the path, with the
    reference's error for anything else (voxelgrids.py:573-578)."""
    if input_path[-3:] == 'xyz':
        return load_xyz(input_path)
    if input_path[-3:] == 'pdb':
        return load_pdb(input_path)
    raise Exception('files must be a .pdb or .xyz file')


def most_common_element(input_path):
    """Most frequent element symbol of a structure file; ties go to the element
    seen first, like collections.Counter.most_common (utilities.py:61-75)."""
    _, elements = load_structure(input_path)
    first_seen, counts = {}, {}
    for i, el in enumerate(elements.tolist()):
        counts[el] = counts.get(el, 0) + 1
        first_seen.setdefault(el, i)
    return max(counts, key=lambda el: (counts[el], -first_seen[el]))


def calc_real_space_abc(a_mag, b_mag, c_mag, alpha_deg, beta_deg, gamma_deg):
    """Cell vectors a (along x), b (in the xy plane), c from lengths and angles
    (utilities.py:276-301); the same NumPy expressions, so the components are
    bit-identical to the reference's."""
    alpha, beta, gamma = np.deg2rad(alpha_deg), np.deg2rad(beta_deg), np.deg2rad(gamma_deg)
    V = a_mag * b_mag * c_mag * np.sqrt(1 - np.cos(alpha) ** 2 - np.cos(beta) ** 2 - np.cos(gamma) ** 2
                                        + 2 * np.cos(alpha) * np.cos(beta) * np.cos(gamma))
    a = np.array([a_mag, 0, 0])
    b = np.array([b_mag * np.cos(gamma), b_mag * np.sin(gamma), 0])
    c = np.array([c_mag * np.cos(beta),
                  c_mag * (np.cos(alpha) - np.cos(beta) * np.cos(gamma)) / (np.sin(gamma)),
                  V / (a_mag * b_mag * np.sin(gamma))])
    return a, b, c


def load_pdb_cell_params(pdb_path):
    """(a, b, c, alpha, beta, gamma) from the CRYST1 record of a PDB file: fixed
    columns 7-15, 16-24, 25-33, 34-40, 41-47, 48-54 (utilities.py:163-195)."""
    with open(pdb_path, "r") as fh:
        for line in fh:
            if line.startswith("CRYST1"):
                cuts = ((6, 15), (15, 24), (24, 33), (33, 40), (40, 47), (47, 54))
                return tuple(float(line[lo:hi].strip()) for lo, hi in cuts)
    raise ValueError("No CRYST1 line found in the PDB file.")


# ---------------------------------------------------------------------------
# key=value configuration files (simulate_GIWAXS.py --config)
# ---------------------------------------------------------------------------
_TRUE = {"true", "t", "yes", "y", "1", "on"}
_FALSE = {"false", "f", "no", "n", "0", "off"}


def str_to_bool(input_value, default=False):
    """'true'/'yes'/'1'/'on'... -> True, 'false'/'no'/'0'/'off'... -> False, anything else ->
    `default` (ValueError when default is None) (utilities.py:10-41)."""
    text = str(input_value).strip().lower()
    if text in _TRUE:
        return True
    if text in _FALSE:
        return False
    if default is not None:
        return default
    raise ValueError(f"Invalid input for boolean conversion: {input_value}")


def parse_config_file(file_path):
    """{key: value-string} of every line containing '=' (split at the first one; the line is
    stripped, keys and values are not) (utilities.py:43-50)."""
    config = {}
    with open(file_path, "r") as fh:
        for line in fh:
            if "=" in line:
                key, value = line.strip().split("=", 1)
                config[key] = value
    return config


def save_config_to_txt(config, save_path):
    """Echo a configuration as key=value lines (utilities.py:52-55)."""
    with open(save_path, "w") as fh:
        for key, value in config.items():
            fh.write(f"{key}={value}\n")
