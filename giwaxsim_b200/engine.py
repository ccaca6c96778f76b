"""Host orchestration of the two hot-path stages on one GPU.

Stage A (`SliceEngine`): atoms -> per-phi projection -> |FFT|^2 -> 3-D binning
(reference: tools/comparison.py:673-788, tools/voxelgrids.py:311-506).
Stage B (`DetectorEngine`): rotated detector planes gathered from the voxel
grid and summed over orientations (tools/comparison.py:790-870,
tools/detector.py:194-300).

PyTorch is used for device buffers, streams and host<->device copies only; all
arithmetic on the path runs in the hand-written kernels behind the C ABI
(giwaxsim_b200/_lib.py).  Every angle-dependent scalar is evaluated here with
the same NumPy expressions the reference uses, so the integer indices derived
on the device are bit-identical to the reference's.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr

import os

# GIWAXS_B200_STAGED=1 forces the unfused kernels (projection, 2-D FFT, binning
# as separate launches) -- used to compare the two paths; both are CUDA.
STAGED_ONLY = os.environ.get("GIWAXS_B200_STAGED", "0") == "1"
# GIWAXS_B200_FULL_RANGE=1 computes min/max of y' over all atoms instead of the
# convex-hull candidates (cross-check of the candidate reduction).
USE_ALL_ATOMS_FOR_RANGE = os.environ.get("GIWAXS_B200_FULL_RANGE", "0") == "1"
# GIWAXS_B200_EXACT_DETECTOR=1 runs the all-fp64 detector kernel instead of the
# fp32-filtered one (both give identical voxel indices).
EXACT_DETECTOR_ONLY = os.environ.get("GIWAXS_B200_EXACT_DETECTOR", "0") == "1"
# GIWAXS_B200_DETECTOR_TMA=1 feeds the affine detector gather with TMA-staged 8^3 voxel bricks instead of
# L1-cached loads (A/B variant of the same kernel; identical result, measured slower - DESIGN.md 4.4).
DETECTOR_TMA = os.environ.get("GIWAXS_B200_DETECTOR_TMA", "0") == "1"

_checked_devices = set()


def resolve_device(device=None):
    """torch.device of the GPU to use; fails loudly when there is none."""
    if device is None:
        if not torch.cuda.is_available():
            # let the library produce its own message (no CPU fallback)
            call("gx_device_check", 0)
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if device.index not in _checked_devices:
        call("gx_device_check", device.index)
        _checked_devices.add(device.index)
    return device


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_staging = {}


def _pinned_staging(nbytes):
    """Persistent page-locked scratch (grown, never shrunk).  Page-locking is slow (~1 GB/s), so
    results are never handed out in freshly pinned memory; they are DMA'd into this buffer and
    copied out."""
    buf = _staging.get("buf")
    if buf is None or buf.numel() < nbytes:
        _staging["buf"] = None
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
        _staging["buf"] = buf
    return buf


PIPELINED_DOWNLOAD_MIN_BYTES = 64 << 20
SHARED_RESULT_MIN_BYTES = 128 << 20        # float64 bytes above which N ranks of a node share one host conversion
PIPELINED_DOWNLOAD_CHUNKS = 8
_host_pool = []        # [{"buf": flat float64 CPU tensor, "live": weakref to the array handed out}]
_HOST_POOL_MAX = 4
_HOST_POOL_MIN_BYTES = 16 << 20       # first touch of a fresh 33 MB detector image costs as much as its DMA
_HOST_POOL_PIN_MAX_BYTES = 96 << 20   # pool entries up to this size are page-locked


def _result_buffer(shape):
    """float64 CPU tensor for a result.  Large results come from a small pool of buffers whose
    pages are already faulted in (first touch of a fresh 0.5 GB array costs more than filling it);
    a buffer is reused only once the array previously handed out on it - and every view of it - has
    been garbage collected, so arrays a caller still holds never change."""
    import weakref
    n = int(np.prod(shape))
    if n * 8 < _HOST_POOL_MIN_BYTES:
        return torch.empty(shape, dtype=torch.float64), None
    for e in _host_pool:
        if e["buf"].numel() == n and (e["live"] is None or e["live"]() is None):
            return e["buf"].view(shape), e
    if len(_host_pool) >= _HOST_POOL_MAX:
        for i, e in enumerate(_host_pool):
            if e["live"] is None or e["live"]() is None:
                _host_pool.pop(i)                       # a free buffer of another size makes room
                break
        else:
            return torch.empty(shape, dtype=torch.float64), None
    # medium results (the detector image) sit in page-locked memory: the fp64 tensor is DMA'd straight into
    # the array the caller gets, no host-side copy (page-locking is paid once per pool entry)
    pin = torch.cuda.is_available() and n * 8 <= _HOST_POOL_PIN_MAX_BYTES
    e = {"buf": torch.empty(n, dtype=torch.float64, pin_memory=pin), "live": None, "pinned": pin}
    _host_pool.append(e)
    return e["buf"].view(shape), e


def to_host_f64(t, out=None, replicated=False):
    """Device tensor -> float64 NumPy array: one DMA of the tensor in its own dtype (fp32 voxel
    grids cross PCIe at half the bytes) into the persistent pinned buffer, then a multi-threaded
    widening copy on the host into a fresh array (or into `out`, a float64 NumPy view).
    replicated=True: every rank holds the same tensor (after an all-reduce) -> the conversion is
    shared between the ranks of the node (parallel.shared_result_f64)."""
    t = t.contiguous()
    if out is None:
        from . import parallel
        # Large results live in a pooled memory file and are handed out as a private copy-on-write mapping:
        # N ranks of a node produce the bytes once (each converts its slab), and the kernel tracks whether
        # the caller ever writes to the array (parallel.result_unmodified), which is what lets the next
        # driver re-use the device copy without re-reading 0.5 GB.  (Small results - the 33 MB detector
        # image - are cheaper to convert per rank.)
        shared = parallel.shared_result_f64(t, lambda dev_slice, view: to_host_f64(dev_slice, out=view),
                                            min_bytes=SHARED_RESULT_MIN_BYTES, single=True)
        if shared is not None:
            return shared
    nbytes = t.numel() * t.element_size()
    entry = None
    if out is None:
        out, entry = _result_buffer(tuple(t.shape))
        if entry is not None and entry.get("pinned"):
            out.copy_(t.to(torch.float64), non_blocking=True)      # widened on the device, one DMA, no host work
            torch.cuda.current_stream().synchronize()
            res = out.numpy()
            import weakref
            entry["live"] = weakref.ref(res)
            return res
    else:
        out = torch.from_numpy(out).view(t.shape)
    stage = _pinned_staging(nbytes)[:nbytes].view(t.dtype).view(t.shape)
    # torchrun pins OMP_NUM_THREADS=1; the widening copy of a large grid is worth a few host
    # threads per rank (never more than the cores this rank can fairly claim)
    before = torch.get_num_threads()
    want = max(1, min(16, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    threaded = nbytes >= (8 << 20) and want > before
    if threaded:
        torch.set_num_threads(want)
    try:
        if nbytes >= PIPELINED_DOWNLOAD_MIN_BYTES and out.is_contiguous():
            # large grids: the DMA is cut into chunks and chunk k is widened on the host while chunk
            # k+1 is still on the wire (PCIe time + one chunk instead of PCIe time + the whole widening)
            src, dst, stg = t.view(-1), out.view(-1), stage.view(-1)
            n = src.numel()
            step = -(-n // PIPELINED_DOWNLOAD_CHUNKS)
            marks = []
            for lo in range(0, n, step):
                hi = min(n, lo + step)
                stg[lo:hi].copy_(src[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                marks.append((lo, hi, ev))
            widen = t.dtype == torch.float32 and dst.dtype == torch.float64
            for lo, hi, ev in marks:
                ev.synchronize()
                if widen:
                    # library pool + non-temporal stores: torch's cast copy reads every destination line first
                    call("gx_host_widen_f32_f64", stg.data_ptr() + 4 * lo, dst.data_ptr() + 8 * lo, hi - lo, want)
                else:
                    dst[lo:hi].copy_(stg[lo:hi])
        else:
            stage.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            out.copy_(stage)
    finally:
        if threaded:
            torch.set_num_threads(before)
    res = out.numpy()
    if entry is not None:
        import weakref
        entry["live"] = weakref.ref(res)                # views of `res` keep it alive (numpy collapses .base)
    return res


def device_checksum(t, widen_f32=False):
    """Wrap-around 64-bit word sum of a device tensor (gx_checksum64) as a Python int; with
    widen_f32 the sum of the float64 bit patterns of an fp32 tensor - what host_checksum()
    returns for the widened host copy."""
    t = t.contiguous()
    out = torch.zeros(1, dtype=torch.int64, device=t.device)
    n = t.numel() if widen_f32 else t.numel() * t.element_size() // 8
    with torch.cuda.device(t.device):
        call("gx_checksum64", ptr(t), int(n), int(bool(widen_f32)), ptr(out), _stream())
        return int(out.item()) & 0xFFFFFFFFFFFFFFFF


def host_checksum(a):
    """The same checksum of a C-contiguous host array whose byte size is a multiple of 8
    (multi-threaded: ~20 ms for a 0.5 GB grid), or None for anything else."""
    if not isinstance(a, np.ndarray) or not a.flags.c_contiguous or a.nbytes % 8:
        return None
    flat = torch.from_numpy(a.reshape(-1).view(np.uint8)).view(torch.int64)
    before = torch.get_num_threads()
    want = max(1, min(16, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    try:
        if a.nbytes >= (8 << 20) and want > before:
            torch.set_num_threads(want)
        return int(flat.sum().item()) & 0xFFFFFFFFFFFFFFFF
    finally:
        torch.set_num_threads(before)


def _dev(a, device, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device, non_blocking=True)


# ---------------------------------------------------------------------------
# species coding
# ---------------------------------------------------------------------------
def encode_values(values, max_species=_lib.GX_MAX_SPECIES):
    """Code an array with few distinct entries as uint8 without sorting it.
    Returns (codes uint8 [A], uniques list) or (None, None) when there are more
    than `max_species` distinct values."""
    values = np.asarray(values)
    A = values.shape[0]
    if values.dtype.kind == "U" and values.dtype.itemsize in (4, 8) and A > 0:
        # element symbols ('<U1' / '<U2'): ASCII code points -> 14-bit key -> lookup table,
        # three vector passes instead of one string comparison pass per element type
        cp = np.ascontiguousarray(values).view(np.uint32).reshape(A, -1)
        if int(cp.max()) < 128:
            key = cp[:, 0] if cp.shape[1] == 1 else cp[:, 0] + (cp[:, 1] << 7)
            present = np.flatnonzero(np.bincount(key, minlength=1 << 14))
            if len(present) > max_species:
                return None, None
            lut = np.zeros(1 << 14, dtype=np.uint8)
            lut[present] = np.arange(len(present), dtype=np.uint8)
            uniques = [values.dtype.type("".join(chr(c) for c in (k & 127, k >> 7) if c)) for k in present]
            return lut[key], uniques
    codes = np.full(A, 255, dtype=np.uint8)
    uniques = []
    todo = np.ones(A, dtype=bool)
    while True:
        first = int(np.argmax(todo))
        if not todo[first]:
            break
        if len(uniques) == max_species:
            return None, None
        v = values[first]
        hit = values == v
        codes[hit] = len(uniques)
        uniques.append(v)
        todo &= ~hit
    return codes, uniques


def encode_elements_device(elements, device, max_species=_lib.GX_MAX_SPECIES):
    """Code a NumPy '<U1' / '<U2' element array as uint8 ON THE DEVICE (the raw code points are
    uploaded; a 10 M-atom array costs three full passes on the host otherwise).
    Returns (codes uint8 device tensor [A], unique symbols, atoms per symbol) or None when the
    array is not of that form / has more than `max_species` distinct symbols."""
    elements = np.asarray(elements)
    A = elements.shape[0] if elements.ndim == 1 else 0
    if A == 0 or elements.dtype.kind != "U" or elements.dtype.itemsize not in (4, 8):
        return None
    width = elements.dtype.itemsize // 4
    from . import parallel
    cp = parallel.upload_replicated(np.ascontiguousarray(elements).view(np.uint32).view(np.int32), device)
    hist = torch.empty(16385, dtype=torch.int32, device=device)
    st = _stream()
    call("gx_species_histogram", ptr(cp), width, A, ptr(hist), st)
    h = hist.cpu().numpy()
    if h[16384] != 0:
        return None
    present = np.flatnonzero(h[:16384])
    if len(present) > max_species:
        return None
    lut = np.zeros(16384, dtype=np.uint8)
    lut[present] = np.arange(len(present), dtype=np.uint8)
    d_lut = _dev(lut, device)
    codes = torch.empty(A, dtype=torch.uint8, device=device)
    call("gx_species_codes", ptr(cp), width, A, ptr(d_lut), ptr(codes), st)
    uniques = [elements.dtype.type("".join(chr(c) for c in (int(k) & 127, int(k) >> 7) if c)) for k in present]
    return codes, uniques, h[present].astype(np.int64)


# ---------------------------------------------------------------------------
# slab builder (next row N3)
# ---------------------------------------------------------------------------
def build_slab(cell_coords, cell_codes, sizes, vectors, device):
    """Tile the unit cell and cut the centred slab on the device (comparison.py:605-671).
    cell_coords [n,3] float64 host, cell_codes [n] uint8 host or None, sizes = (x, y, z) slab
    edge lengths, vectors = (a, b, c) cell vectors.  Returns (coords [M,3] float64 device tensor,
    codes [M] uint8 device tensor or None), atoms in the reference's order, coordinates
    bit-identical to the reference's."""
    a_vec, b_vec, c_vec = (np.asarray(v, dtype=np.float64) for v in vectors)
    x_size, y_size, z_size = sizes
    num = (int(np.ceil(2 * x_size / a_vec[0])), int(np.ceil(2 * y_size / b_vec[1])),
           int(np.ceil(2 * z_size / c_vec[2])))                                   # comparison.py:608-610
    if min(num) < 0:
        raise ValueError("slab sizes and cell vectors must be positive")
    cell = np.ascontiguousarray(cell_coords, dtype=np.float64)
    st = _stream()
    d_cell = _dev(cell, device)
    d_codes = _dev(np.ascontiguousarray(cell_codes, dtype=np.uint8), device) if cell_codes is not None else None
    args = _lib.SlabArgs()
    args.d_cell_xyz, args.d_cell_species = d_cell.data_ptr(), (d_codes.data_ptr() if d_codes is not None else None)
    args.n_cell = int(cell.shape[0])
    args.nx, args.ny, args.nz = num[0] + 1, num[1] + 1, num[2] + 1
    args.ax, args.bx, args.by = float(a_vec[0]), float(b_vec[0]), float(b_vec[1])
    args.cx, args.cy, args.cz = float(c_vec[0]), float(c_vec[1]), float(c_vec[2])
    ref = ctypes.byref(args)
    mm = torch.empty(6, dtype=torch.float64, device=device)
    call("gx_slab_minmax", ref, ptr(mm), st)
    mm = mm.cpu().numpy()
    mins = np.ascontiguousarray(mm[0::2])
    ext = mm[1::2] - mm[0::2]                                                     # comparison.py:639-641
    assert ext[0] > x_size, "x_max must be greater than x_size"
    assert ext[1] > y_size, "y_max must be greater than y_size"
    assert ext[2] > z_size, "z_max must be greater than z_size"
    lo = np.ascontiguousarray((ext - np.array([x_size, y_size, z_size], dtype=np.float64)) / 2)
    hi = np.ascontiguousarray(ext - lo)
    tiles = int(_lib.cdll().gx_slab_tiles(ref))
    offsets = torch.empty(tiles, dtype=torch.int64, device=device)
    total = torch.empty(1, dtype=torch.int64, device=device)
    kept_min = torch.empty(3, dtype=torch.float64, device=device)
    call("gx_slab_count", ref, ptr(mins), ptr(lo), ptr(hi), ptr(offsets), ptr(total), ptr(kept_min), st)
    M = int(total.item())
    if M == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    kmin = np.ascontiguousarray(kept_min.cpu().numpy())
    out = torch.empty((M, 3), dtype=torch.float64, device=device)
    out_codes = torch.empty(M, dtype=torch.uint8, device=device) if d_codes is not None else None
    call("gx_slab_write", ref, ptr(mins), ptr(lo), ptr(hi), ptr(kmin), ptr(offsets), ptr(out), ptr(out_codes), st)
    return out, out_codes


# ---------------------------------------------------------------------------
# stage A
# ---------------------------------------------------------------------------
class FftPlan:
    """Device copy of the twiddle / chirp table for one transform length."""

    _cache = {}

    def __init__(self, N, device):
        nbytes = call("gx_fft_plan_bytes", int(N))
        if nbytes < 0:
            raise _lib.GxError(int(nbytes), _lib.last_error())
        host = np.zeros(max(nbytes // 4, 2), dtype=np.float32)
        call("gx_fft_plan_fill", int(N), ptr(host))
        self.N = int(N)
        self.table = _dev(host, device)

    @classmethod
    def get(cls, N, device):
        key = (int(N), device.index)
        if key not in cls._cache:
            cls._cache[key] = cls(N, device)
        return cls._cache[key]


def check_grid_size(grid_size):
    """Raise a clear error for a real-space grid side the row / column transforms do not cover
    (grid_size = ceil(2 pi / (q_voxel_size r_voxel_size)), comparison.py:710, is unconstrained in
    the reference); called before any upload."""
    if int(call("gx_fft_plan_bytes", int(grid_size))) < 0:
        raise _lib.GxError(_lib.GX_ERR_UNSUPPORTED,
                           "real-space grid of %d^2 pixels is not supported (%s); choose q_voxel_size / "
                           "r_voxel_size so that ceil(2 pi / (q r)) is within the supported sizes"
                           % (grid_size, _lib.last_error()))


def stage_a_geometry(bounds, r_voxel_size, q_voxel_size, max_q):
    """Scalar set-up of voxelgridmaker_fitting (tools/comparison.py:705-731),
    same NumPy expressions.  bounds = (x_bound, y_bound, z_bound)."""
    max_q_diag = np.sqrt(2) * max_q
    if max_q_diag > 2 * np.pi / r_voxel_size:
        raise Exception('Max_q is non-physical for given voxel size')
    grid_size = int(np.ceil(2 * np.pi / (q_voxel_size * r_voxel_size)))
    if grid_size * r_voxel_size < np.min(bounds):
        raise Exception('Calculated real-space bounds smaller than simulation. Please lower delta_q value')
    max_q_diag = max_q_diag + max_q_diag % q_voxel_size
    q_num = ((2 * max_q_diag / q_voxel_size) + 1).astype(int)
    if q_num % 2 == 0:
        q_num += 1
    q_axis = np.linspace(-max_q_diag, max_q_diag, q_num)
    delta_phi_rad = np.arctan(q_voxel_size / max_q_diag)
    phi_num = np.ceil(2 * np.pi / delta_phi_rad).astype(int)
    phis = np.linspace(0, 180 - (180 / phi_num), num=phi_num)
    return grid_size, int(q_num), q_axis, phis


CHORD_DTYPE = np.dtype([(n, np.float64) for n in
                        ("hor", "ver", "stop1", "stop2", "stop12", "mid", "vcos", "rise",
                         "tan_phi", "tan_theta", "cos_phi", "cos_theta")] +
                       [("mode", np.int32), ("pad", np.int32)])
assert CHORD_DTYPE.itemsize == ctypes.sizeof(_lib.Chord)


def chord_constants(phis, hor_length, ver_length):
    """Per-rotation constants of rectangular_collapse_lengths
    (tools/voxelgrids.py:253-285), called there as (x_vals, y_bound, x_bound, phi).
    Vectorised over the rotations with the same element-wise NumPy operations
    the reference applies to scalars; returns a structured array laid out as
    gx_chord."""
    phis = np.asarray(phis, dtype=np.float64)
    out = np.zeros(len(phis), dtype=CHORD_DTYPE)
    swap = phis > 90
    phi = np.where(swap, phis - 90, phis)
    hor = np.where(swap, np.float64(ver_length), np.float64(hor_length))
    ver = np.where(swap, np.float64(hor_length), np.float64(ver_length))
    theta_rad = np.deg2rad(90 - phi)
    phi_rad = np.deg2rad(phi)
    out["hor"], out["ver"] = hor, ver
    out["mode"] = np.where(phi == 0, 0, np.where(phi == 90, 1, 2))
    with np.errstate(all="ignore"):
        vcos = ver * np.cos(theta_rad)
        s1 = vcos
        s2 = hor * np.cos(phi_rad)
        first = s1 < s2
        out["mid"] = np.where(first, ver / np.sin(theta_rad), hor / np.sin(phi_rad))
        out["stop1"] = np.where(first, s1, s2)
        out["stop2"] = np.where(first, s2, s1)
        out["stop12"] = out["stop1"] + out["stop2"]
        out["vcos"] = vcos
        out["rise"] = np.sqrt(ver ** 2 - vcos ** 2)
        out["tan_phi"], out["tan_theta"] = np.tan(phi_rad), np.tan(theta_rad)
        out["cos_phi"], out["cos_theta"] = np.cos(phi_rad), np.cos(theta_rad)
    return out


def gaussian_weights(sigma):
    """scipy.ndimage._gaussian_kernel1d(sigma, 0, int(4*sigma+0.5))."""
    radius = int(4.0 * float(sigma) + 0.5)
    sigma2 = float(sigma) * float(sigma)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / sigma2 * x ** 2)
    return w / w.sum(), radius


class AtomSet:
    """Device-resident slab: atoms sorted by z pixel row (phi-invariant)."""

    def __init__(self, coords, r_voxel_size, grid_size, device, species=None, table=None, f_values=None):
        """coords: host [A,3] array, or a float64 device tensor (a slab built by build_slab)."""
        on_device = isinstance(coords, torch.Tensor)
        if not on_device:
            coords = np.ascontiguousarray(coords, dtype=np.float64)
        if coords.ndim != 2 or coords.shape[1] != 3 or coords.shape[0] == 0:
            raise ValueError("coords must be a non-empty [A,3] array")
        self.device = device
        self.A = int(coords.shape[0])
        self.N = int(grid_size)
        self.r = float(r_voxel_size)
        st = _stream()
        from . import parallel
        d_coords = coords.contiguous() if on_device else parallel.upload_replicated(coords, device)
        mm = torch.empty(6, dtype=torch.float64, device=device)
        call("gx_coords_minmax", ptr(d_coords), self.A, ptr(mm), st)
        self.minmax = mm.cpu().numpy()
        self.bounds = tuple(self.minmax[2 * k + 1] - self.minmax[2 * k] for k in range(3))
        self.n_species = 0
        d_species = d_f = None
        self.species = self.f = self.table = None
        if species is not None:
            self.n_species = len(table)
            d_species = species if isinstance(species, torch.Tensor) else _dev(species, device)
            self.species = torch.empty(self.A, dtype=torch.uint8, device=device)
            self.table_c128 = np.asarray(table, dtype=np.complex128).copy()
            self.table_host = self.table_c128.astype(np.complex64).view(np.float32).copy()
            self.table = _dev(self.table_host, device)
        else:
            f_host = np.asarray(f_values, dtype=np.complex128)
            self.max_abs_f = (float(np.abs(f_host.real).max()), float(np.abs(f_host.imag).max())) if f_host.size else (0.0, 0.0)
            d_f = _dev(f_host.astype(np.complex64).view(np.float32), device)
            self.f = torch.empty(2 * self.A, dtype=torch.float32, device=device)
        self.xs = torch.empty(self.A, dtype=torch.float64, device=device)
        self.ys = torch.empty(self.A, dtype=torch.float64, device=device)
        self.perm = torch.empty(self.A, dtype=torch.int32, device=device)
        self.row_start = torch.empty(self.N + 2, dtype=torch.int32, device=device)
        cursor = torch.empty(self.N + 2, dtype=torch.int32, device=device)
        call("gx_atoms_sort_rows", ptr(d_coords), self.A, float(self.minmax[4]), self.r, self.N,
             ptr(d_species), ptr(d_f), ptr(self.xs), ptr(self.ys), ptr(self.perm),
             ptr(self.species), ptr(self.f), ptr(self.row_start), ptr(cursor), st)
        self._cand = None
        self._max_row_atoms = None
        self._max_row_abs_f = None

    @property
    def max_row_atoms(self):
        """Most atoms in one z pixel row (sizes the fixed-point scale of the fused row kernel)."""
        if self._max_row_atoms is None:
            rs = self.row_start[:self.N + 1]
            self._max_row_atoms = int((rs[1:] - rs[:-1]).max().item())
        return self._max_row_atoms

    @property
    def max_row_abs_f(self):
        """(max over z rows of sum |Re f|, same for |Im f|): what one pixel of a row can receive at
        most; sizes the fixed-point scale of the fused row kernel (gx_fused_args.max_row_abs_*)."""
        if self._max_row_abs_f is None:
            out = torch.empty(2, dtype=torch.float64, device=self.device)
            tab = None
            if self.n_species:
                tab = np.ascontiguousarray(np.abs(np.stack([self.table_c128.real, self.table_c128.imag], axis=1)))
            with torch.cuda.device(self.device):
                call("gx_row_abs_f_max", ptr(self.species), ptr(self.f), ptr(self.row_start), self.N, ptr(tab),
                     self.n_species, ptr(out), _stream())
                m = out.cpu().numpy() * (1.0 + 1e-9)
            self._max_row_abs_f = (float(m[0]), float(m[1]))
        return self._max_row_abs_f

    def candidates(self):
        """(xs, ys, count) of the atoms that can be extreme in y' (built lazily, once)."""
        if self._cand is None:
            self._cand = (self.xs, self.ys, self.A) if USE_ALL_ATOMS_FOR_RANGE else extreme_candidates(self)
        return self._cand


def convex_polygon(points):
    """Counter-clockwise convex hull (Andrew's monotone chain) of a few 2-D points."""
    pts = sorted(set(map(tuple, np.asarray(points, dtype=np.float64))))
    if len(pts) < 3:
        return np.asarray(pts)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.asarray(lower[:-1] + upper[:-1])


def extreme_candidates(atoms, n_dir=32, capacity=1 << 20):
    """Subset of atoms that can attain min/max of y' for some rotation: those not
    strictly inside (by eps) the polygon spanned by the atoms that are extreme
    along 2*n_dir directions.  Returns (xs, ys, count) device tensors; falls back
    to all atoms for degenerate (collinear) slabs."""
    dev, A = atoms.device, atoms.A
    st = _stream()
    theta = np.linspace(0.0, np.pi, n_dir, endpoint=False)
    d_sn, d_cs = _dev(np.sin(theta), dev), _dev(np.cos(theta), dev)
    yr = torch.empty(2 * n_dir, dtype=torch.float64, device=dev)
    call("gx_slice_yrange", ptr(atoms.xs), ptr(atoms.ys), A, ptr(d_sn), ptr(d_cs), n_dir, ptr(yr), st)
    idx = torch.empty(2 * n_dir, dtype=torch.int32, device=dev)
    call("gx_extreme_atoms", ptr(atoms.xs), ptr(atoms.ys), A, ptr(d_sn), ptr(d_cs), ptr(yr), n_dir, ptr(idx), st)
    sel = torch.unique(idx.to(torch.int64))
    sel = sel[sel < A]
    pts = torch.stack([atoms.xs[sel], atoms.ys[sel]], dim=1).cpu().numpy()
    poly = convex_polygon(pts)
    if len(poly) < 3 or len(poly) > 64:
        return atoms.xs, atoms.ys, A
    nxt = np.roll(poly, -1, axis=0)
    e = nxt - poly
    length = np.hypot(e[:, 0], e[:, 1])
    normal = np.stack([-e[:, 1], e[:, 0]], axis=1) / length[:, None]          # inward for CCW order
    edges = np.concatenate([normal, -(normal * poly).sum(axis=1, keepdims=True)], axis=1)
    eps = 1e-7 * (1.0 + float(np.abs(atoms.minmax[:4]).max()))
    cap = int(min(A, capacity))
    out_x = torch.empty(cap, dtype=torch.float64, device=dev)
    out_y = torch.empty(cap, dtype=torch.float64, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    d_edges = _dev(np.ascontiguousarray(edges), dev)
    call("gx_hull_filter", ptr(atoms.xs), ptr(atoms.ys), A, ptr(d_edges), int(len(poly)), eps, ptr(count),
         ptr(out_x), ptr(out_y), cap, st)
    n = int(count.item())
    if n > cap or n == 0:
        return atoms.xs, atoms.ys, A
    return out_x[:n].contiguous(), out_y[:n].contiguous(), n


class SliceEngine:
    """Runs phi slices of one slab into sum/count accumulators on one GPU."""

    def __init__(self, coords, r_voxel_size, q_axis, grid_size, avg_voxel_f, x_bound, y_bound,
                 fill_bkg, smooth, species=None, table=None, f_values=None, device=None,
                 count3d=False, accumulators=None, atoms=None, window=None):
        """window=(lo, hi): accumulate only the voxels lo <= i < hi of every axis (the crop of
        downselect_voxelgrid commutes with the sum) into [hi-lo]^3 grids; columns and rows that
        fall outside are neither transformed nor binned."""
        self.device = resolve_device(device)
        with torch.cuda.device(self.device):
            self.N = int(grid_size)
            self.r = float(r_voxel_size)
            self.atoms = atoms if atoms is not None else AtomSet(
                coords, r_voxel_size, grid_size, self.device, species, table, f_values)
            self.plan = FftPlan.get(self.N, self.device)
            self.fill_bkg = bool(fill_bkg)
            self.sigma = int(smooth) if smooth else 0
            self.x_bound, self.y_bound = x_bound, y_bound
            self.avg_voxel_f = complex(avg_voxel_f)
            # voxelgrids.py:344 and :358 / :365
            self.max_voxels = np.sqrt(x_bound ** 2 + y_bound ** 2) // r_voxel_size
            ped = avg_voxel_f * self.max_voxels
            self.has_pedestal = self.fill_bkg or self.sigma > 0
            self.pedestal = complex(ped) if self.has_pedestal else 0j
            self.q_axis = np.asarray(q_axis, dtype=np.float64)
            self.q_num = int(self.q_axis.shape[0])
            self.qmin, self.qmax = float(np.min(self.q_axis)), float(np.max(self.q_axis))
            self.dq = float(np.diff(self.q_axis)[0])
            # voxelgrids.py:382-385 (phi-invariant row axis) and its bin indices
            self.q_fft = np.fft.fftshift(np.fft.fftfreq(self.N, d=r_voxel_size) * 2 * np.pi)
            self.q_fft_max = np.max(self.q_fft)
            dev = self.device
            st = _stream()
            self.row_index = torch.empty(self.N, dtype=torch.int32, device=dev)
            self.d_q_fft = _dev(self.q_fft, dev)
            call("gx_axis_row_index", ptr(self.d_q_fft), self.N, self.qmin, self.qmax, self.dq,
                 self.q_num, ptr(self.row_index), st)
            self.window = None if window is None else (int(window[0]), int(window[1]))
            self.q_out = self.q_num            # side of the accumulator grids
            if self.window is not None:
                lo, hi = self.window
                call("gx_window_indices", ptr(self.row_index), self.N, self.q_num, lo, hi, 0, st)
                self.q_out = hi - lo
            self.vsum_store, self.vsum_is_partial = None, False
            if accumulators is not None:
                self.vsum, self.count3, self.count2 = accumulators
            else:
                # padded to whole-column slabs per rank, so that N ranks can reduce-scatter it in place
                from . import parallel
                _, padded = parallel.padded_columns(self.q_out, parallel.rank_world()[1])
                self.vsum_store = torch.zeros(padded, dtype=torch.float32, device=dev)
                self.vsum = self.vsum_store[:self.q_out ** 3]
                self.count3 = torch.zeros(self.q_out ** 3, dtype=torch.int32, device=dev) if count3d else None
                self.count2 = None if count3d else torch.zeros(self.q_out ** 2, dtype=torch.int32, device=dev)
            self.dc = torch.zeros(16, dtype=torch.float64, device=dev)    # fp64 side sums of the DC samples (gx_fold_dc)
            self.row_hist = torch.zeros(self.q_out, dtype=torch.int32, device=dev)
            call("gx_row_histogram", ptr(self.row_index), self.N, self.q_out, ptr(self.row_hist), st)
            if self.sigma > 0:
                w, self.gauss_radius = gaussian_weights(self.sigma)
                self.gauss = _dev(w, dev)
            else:
                self.gauss, self.gauss_radius = None, 0
            self.slices_done = 0
            # kept shifted rows form one interval (qz is monotone); kept columns of a
            # slice are spaced 2*qmax_fft/(N-1) apart along a line through the voxel box
            ri = self.row_index.cpu().numpy()
            kept = np.where(ri >= 0)[0]
            self.row_lo, self.row_hi = (int(kept[0]), int(kept[-1]) + 1) if len(kept) else (0, 0)
            step = 2.0 * self.q_fft_max / (self.N - 1)
            reach = max(abs(self.qmin), abs(self.qmax))
            if self.window is not None:
                # columns are kept when both q components fall in [axis[lo], axis[hi-1] + dq)
                reach = max(abs(self.q_axis[self.window[0]]), abs(self.q_axis[self.window[1] - 1] + self.dq))
            self.KC = int(min(self.N, (int(2.0 * np.sqrt(2.0) * reach / step) + 4 + 7) // 8 * 8))

    # -- per-batch host scalars -------------------------------------------
    def _phi_scalars(self, phis):
        phis = np.asarray(phis, dtype=np.float64)
        phi_rad = np.radians(phis)                       # utilities.py:305
        sn, cs = np.sin(phi_rad), np.cos(phi_rad)
        neg = np.deg2rad(-phis)                          # voxelgrids.py:396-399
        right_qy = self.q_fft_max * np.cos(neg)
        right_qx = -self.q_fft_max * np.sin(neg)
        return sn, cs, -right_qx, right_qx, -right_qy, right_qy

    def batch_size(self, budget_bytes=4 << 30):
        per = 20 * self.N * self.N
        return int(max(1, min(64, budget_bytes // per)))

    def prepare(self, phis):
        """Device-side per-rotation tables for a batch: y range, bbox, row vectors,
        column indices.  Returns a dict of tensors (kept alive by the caller)."""
        dev, N, n = self.device, self.N, len(phis)
        a = self.atoms
        st = _stream()
        sn, cs, xl, xr, yl, yr = self._phi_scalars(phis)
        t = dict(n=n)
        t["sin"], t["cos"] = _dev(sn, dev), _dev(cs, dev)
        t["yrange"] = torch.empty(2 * n, dtype=torch.float64, device=dev)
        cx, cy, cn = a.candidates()
        call("gx_slice_yrange", ptr(cx), ptr(cy), cn, ptr(t["sin"]), ptr(t["cos"]), n, ptr(t["yrange"]), st)
        t["bbox"] = torch.empty(4 * n, dtype=torch.int32, device=dev)
        scratch = torch.empty(n + 1, dtype=torch.int32, device=dev)
        call("gx_slice_bbox", ptr(a.xs), ptr(a.ys), ptr(a.row_start), N, self.r, ptr(t["sin"]), ptr(t["cos"]),
             ptr(t["yrange"]), n, ptr(t["bbox"]), ptr(scratch), st)
        t["base"] = torch.empty(2 * n * N, dtype=torch.float32, device=dev)
        d_chord = None
        if self.fill_bkg:
            ch = chord_constants(phis, self.y_bound, self.x_bound)
            t["chord"] = _dev(ch.view(np.uint8), dev)
            d_chord = t["chord"]
        # blend mask x "inside the atom box" indicator per column / per row
        t["my"] = torch.empty(n * N, dtype=torch.float32, device=dev)
        t["mz"] = torch.empty(n * N, dtype=torch.float32, device=dev)
        t["dmy"] = torch.empty(2 * n * N, dtype=torch.float32, device=dev)
        call("gx_slice_vectors", ptr(d_chord), ptr(t["bbox"]), n, N, self.r, float(self.max_voxels),
             self.avg_voxel_f.real, self.avg_voxel_f.imag, self.pedestal.real, self.pedestal.imag,
             int(self.fill_bkg), self.sigma, ptr(self.gauss), self.gauss_radius,
             ptr(t["base"]), ptr(t.get("my")), ptr(t.get("mz")), ptr(t["dmy"]), st)
        t["col"] = torch.empty(n * N, dtype=torch.int32, device=dev)
        # keep the four end-point arrays referenced until the launch: temporaries
        # would be recycled by the caching allocator and alias each other
        t["ends"] = [_dev(a, dev) for a in (xl, xr, yl, yr)]
        call("gx_slice_col_index", ptr(t["ends"][0]), ptr(t["ends"][1]), ptr(t["ends"][2]), ptr(t["ends"][3]),
             n, N, self.qmin, self.qmax, self.dq, self.q_num, ptr(t["col"]), st)
        if self.window is not None:
            call("gx_window_indices", ptr(t["col"]), n * N, self.q_num, self.window[0], self.window[1], 1, st)
        # first / one-past-last kept column of every rotation (one launch for the whole run)
        t["colrange"] = torch.empty(2 * n, dtype=torch.int32, device=dev)
        call("gx_slice_col_range", ptr(t["col"]), n, N, ptr(t["colrange"]), st)
        return t

    def project(self, t, grid):
        a = self.atoms
        call("gx_project_slices", ptr(a.xs), ptr(a.ys), ptr(a.species), ptr(a.f), ptr(a.row_start),
             ptr(a.table), a.n_species, ptr(t["sin"]), ptr(t["cos"]), ptr(t["yrange"]), ptr(t["bbox"]),
             ptr(t["base"]), ptr(t.get("my")), ptr(t.get("mz")), t["n"], self.N, self.r,
             self.pedestal.real, self.pedestal.imag, int(self.fill_bkg), self.sigma, ptr(grid), _stream())

    def fft(self, grid, work, iq2d, n):
        call("gx_fft2_abs2_shift", ptr(grid), ptr(work), ptr(iq2d), n, self.N, ptr(self.plan.table),
             0.0, 0.0, _stream())

    def bin(self, t, iq2d):
        call("gx_bin_slices", ptr(iq2d), t["n"], self.N, self.N, ptr(t["col"]), self.N, ptr(self.row_index),
             self.q_out, ptr(self.vsum), ptr(self.count3), ptr(self.count2), _stream())

    def check_bbox(self, t):
        bb = t["bbox"].cpu().numpy().reshape(-1, 4)
        if (bb[:, 1] < 0).any():
            # the reference's np.min on an empty selection (voxelgrids.py:346)
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        return bb

    def fused_batch_size(self, budget_bytes=2 << 30):
        cap = int(os.environ.get("GIWAXS_B200_FUSED_BATCH", "128"))     # what the constant tables of the row kernel hold
        return int(max(1, min(cap, budget_bytes // (8 * self.N * self.KC))))

    def fused(self, t, work):
        """F1 + F2 for one prepared batch (gx_slices_fused)."""
        a = self.atoms
        n = t["n"]
        args = _lib.FusedArgs()
        for name, tensor in (("d_xs", a.xs), ("d_ys", a.ys), ("d_species", a.species), ("d_f", a.f),
                             ("d_row_start", a.row_start), ("d_table", a.table), ("d_sin", t["sin"]),
                             ("d_cos", t["cos"]), ("d_yrange", t["yrange"]), ("d_bbox", t["bbox"]),
                             ("d_dmy", t["dmy"]), ("d_mz", t.get("mz")),
                             ("d_plan", self.plan.table), ("d_col", t["col"]), ("d_colrange", t["colrange"]),
                             ("d_row_index", self.row_index), ("d_work", work), ("d_sum", self.vsum),
                             ("d_count2", self.count2), ("d_dc", self.dc)):
            setattr(args, name, None if tensor is None else tensor.data_ptr())
        args.r = self.r
        args.pedestal_re, args.pedestal_im = self.pedestal.real, self.pedestal.imag
        args.avg_f_re, args.avg_f_im = self.avg_voxel_f.real, self.avg_voxel_f.imag
        if a.n_species:
            for k, v in enumerate(a.table_c128):
                args.table_f64[2 * k], args.table_f64[2 * k + 1] = float(v.real), float(v.imag)
        else:
            args.max_abs_f_re, args.max_abs_f_im = a.max_abs_f
        args.max_row_atoms = a.max_row_atoms
        args.max_row_abs_re, args.max_row_abs_im = a.max_row_abs_f
        args.n_species, args.n_phi, args.N, args.KC, args.q_num = a.n_species, n, self.N, self.KC, self.q_out
        args.row_lo, args.row_hi = self.row_lo, self.row_hi
        args.fill_bkg, args.smooth_sigma = int(self.fill_bkg), self.sigma
        if self.timers is None:
            call("gx_slices_fused", ctypes.byref(args), _stream())
        else:
            # per-kernel CUDA-event timing (bench.py): the two launches are issued by two calls
            for name, phase in (("rows", 1), ("cols", 2)):
                args.phases = phase
                self._timed(name, call, "gx_slices_fused", ctypes.byref(args), _stream())

    def run_fused(self, phis):
        """Production path: two fused launches per batch, nothing N x N in HBM."""
        phis = np.asarray(phis, dtype=np.float64)
        if self.count2 is None:
            raise ValueError("the fused path accumulates rank-1 counts (count3d=False)")
        with torch.cuda.device(self.device):
            B = self.fused_batch_size()
            if len(phis) > B:
                # equal batches (1800 rotations: 29 x 63 instead of 28 x 64 + 8): no short last launch pair
                B = -(-len(phis) // -(-len(phis) // B))
            N = self.N
            # rows outside the atom band are never written; the TMA-fed column kernel reads every row slot, so
            # the buffer is zero-filled once and kept for the engine's later runs (same atoms, same band)
            need = min(B, len(phis)) * N * self.KC * 2
            work = getattr(self, "_work", None)
            if work is None or work.numel() < need:
                alloc = torch.zeros if call("gx_fused_wants_zeroed_work", N, self.KC) else torch.empty
                work = self._work = alloc(need, dtype=torch.float32, device=self.device)
            # per-rotation tables for the whole run in one set of launches, then one
            # pair of fused launches per batch on views of them
            full = self._timed("prepare", self.prepare, phis)
            per_phi = {"sin": 1, "cos": 1, "yrange": 2, "bbox": 4, "base": 2 * N, "my": N, "mz": N, "dmy": 2 * N, "col": N,
                       "colrange": 2}
            for i0 in range(0, len(phis), B):
                n = min(B, len(phis) - i0)
                t = {k: full[k][i0 * w:(i0 + n) * w] for k, w in per_phi.items()}
                t["n"] = n
                self.fused(t, work)
                self.slices_done += n
            call("gx_fold_dc", ptr(self.dc), ptr(self.vsum), _stream())
            torch.cuda.current_stream().synchronize()
            self.check_bbox(full)
            cr = full["colrange"].cpu().numpy().reshape(-1, 2)
            self.mean_kept_columns = float(np.maximum(cr[:, 1] - cr[:, 0], 0).mean())

    def run(self, phis, capture=None, staged=None):
        """Accumulate the given phi slices.  capture: optional dict receiving
        host copies of the per-slice intermediates (parity probes; forces the
        staged kernels, which materialise the pre-FFT grid and |FFT|^2 image)."""
        phis = np.asarray(phis, dtype=np.float64)
        if len(phis) == 0:
            return
        if staged is None:
            staged = capture is not None or self.count2 is None or STAGED_ONLY
        if not staged:
            return self.run_fused(phis)
        with torch.cuda.device(self.device):
            B = self.batch_size()
            N, dev = self.N, self.device
            nb = min(B, len(phis))
            grid = torch.empty(nb * N * N * 2, dtype=torch.float32, device=dev)
            work = torch.empty_like(grid)
            iq2d = torch.empty(nb * N * N, dtype=torch.float32, device=dev)
            boxes = []
            for i0 in range(0, len(phis), B):
                chunk = phis[i0:i0 + B]
                t = self._timed("prepare", self.prepare, chunk)
                boxes.append(t["bbox"])
                self._timed("project", self.project, t, grid)
                self._timed("fft2", self.fft, grid, work, iq2d, t["n"])
                self._timed("bin", self.bin, t, iq2d)
                if capture is not None:
                    n = t["n"]
                    capture.setdefault("bbox", []).append(t["bbox"].cpu().numpy().reshape(n, 4))
                    capture.setdefault("yrange", []).append(t["yrange"].cpu().numpy().reshape(n, 2))
                    capture.setdefault("col", []).append(t["col"].cpu().numpy().reshape(n, N))
                    if capture.get("want_grids"):
                        g = grid[:n * N * N * 2].cpu().numpy().view(np.complex64).reshape(n, N, N)
                        capture.setdefault("grid", []).append(g.copy())
                        capture.setdefault("iq_2d", []).append(iq2d[:n * N * N].cpu().numpy().reshape(n, N, N).copy())
                self.slices_done += len(chunk)
            torch.cuda.current_stream().synchronize()
            # one deferred host check for the whole run (keeps the launch queue full)
            self.check_bbox({"bbox": torch.cat(boxes)})

    timers = None

    def _timed(self, name, fn, *args):
        """Run fn; when self.timers is a dict, bracket it with CUDA events on the
        launching stream (bench.py reads the per-kernel totals)."""
        if self.timers is None:
            return fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*args)
        e1.record()
        self.timers.setdefault(name, []).append((e0, e1))
        return out

    def collect_timers(self):
        """{kernel group: total ms} from the recorded event pairs."""
        torch.cuda.synchronize()
        return {k: float(sum(a.elapsed_time(b) for a, b in v)) for k, v in (self.timers or {}).items()}

    def atom_indices(self, phi):
        """(y_idx, z_idx) int64 of every atom in original order for one phi (probe)."""
        with torch.cuda.device(self.device):
            t = self.prepare(np.array([phi], dtype=np.float64))
            a = self.atoms
            sn, cs, *_ = self._phi_scalars(np.array([phi], dtype=np.float64))
            shift = float(t["yrange"][0].item())
            y = torch.empty(a.A, dtype=torch.int64, device=self.device)
            z = torch.empty(a.A, dtype=torch.int64, device=self.device)
            call("gx_atom_pixel_indices", ptr(a.xs), ptr(a.ys), ptr(a.perm), ptr(a.row_start), a.A, self.N,
                 self.r, float(sn[0]), float(cs[0]), shift, ptr(y), ptr(z), _stream())
            return y.cpu().numpy(), z.cpu().numpy(), t["bbox"].cpu().numpy()

    def counts(self):
        """Per-voxel sample counts as an int64 host array [q,q,q] (q = window size if windowed)."""
        q = self.q_out
        if self.count3 is not None:
            return self.count3.cpu().numpy().astype(np.int64).reshape(q, q, q)
        # rank-1 form: every slice shares the row table, so
        # count[iy,ix,iz] = (kept columns that hit (iy,ix), all slices) * (rows that hit iz)
        h = self.count2.cpu().numpy().astype(np.int64).reshape(q, q)
        m = self.row_hist.cpu().numpy().astype(np.int64)
        return h[:, :, None] * m[None, None, :]

    def sums(self):
        q = self.q_out
        return self.vsum.cpu().numpy().reshape(q, q, q)


CARBON_Z = 6.0
CARBON_AFF = (2.31, 20.8439, 1.02, 10.2075, 1.5886, 0.5687, 0.865, 51.6512, 0.2156)


def crop_range(axis, max_val):
    """downselect_voxelgrid's index range (tools/voxelgrids.py:36-46)."""
    lim = max_val + np.abs(axis[1] - axis[0])
    idx = np.where(np.abs(axis) < lim)[0]
    return int(idx[0]), int(idx[-1]) + 1


def finalize_voxels(vsum, count3, count2, row_hist, q_axis, max_q, device, window=None, crop=True, f0=True,
                    out=None, columns=(0, -1), sync=True):
    """sum/count, crop, carbon f0 weighting -> (iq fp32 device [V,V,V], axis).
    window: the accumulators already cover only that index window (SliceEngine(window=...)).
    crop=False keeps the whole axis and f0=False skips the weighting: the plain grid
    generate_voxel_grid_low_mem returns (voxelgrids.py:633-641)."""
    q_num = int(q_axis.shape[0])
    lo, hi = crop_range(q_axis, max_q) if crop else (0, q_num)
    V = hi - lo
    with torch.cuda.device(device):
        iq = out if out is not None else torch.empty(V * V * V, dtype=torch.float32, device=device)
        aff = np.asarray(CARBON_AFF, dtype=np.float64) if f0 else None
        if window is not None:
            if tuple(window) != (lo, hi):
                raise ValueError("accumulator window %s is not the crop range %s" % (tuple(window), (lo, hi)))
            d_axis = _dev(q_axis[lo:hi], device)
            call("gx_voxel_finalize", ptr(vsum), ptr(count3), ptr(count2), ptr(row_hist), V, 0, V,
                 ptr(d_axis), ptr(aff), CARBON_Z, int(columns[0]), int(columns[1]), ptr(iq), _stream())
        else:
            d_axis = _dev(q_axis, device)
            call("gx_voxel_finalize", ptr(vsum), ptr(count3), ptr(count2), ptr(row_hist), q_num, lo, hi,
                 ptr(d_axis), ptr(aff), CARBON_Z, int(columns[0]), int(columns[1]), ptr(iq), _stream())
        if sync:
            torch.cuda.current_stream().synchronize()
    return iq[:V * V * V].view(V, V, V), q_axis[lo:hi].copy()


def scale_shell(iq, axis, lower, upper, factor, device):
    """iq[qy,qx,qz] *= factor for lower < |q| <= upper, in place on the device
    (shell mask of voxelgrids.py:650,706-707)."""
    V = int(iq.shape[0])
    with torch.cuda.device(device):
        d_axis = _dev(np.asarray(axis, dtype=np.float64), device)
        call("gx_voxel_shell_scale", ptr(iq), V, ptr(d_axis), float(lower), float(upper), float(factor), _stream())
        torch.cuda.current_stream().synchronize()


# ---------------------------------------------------------------------------
# stage B
# ---------------------------------------------------------------------------
def orientation_tables(corners, psis, psi_w, phis, phi_w, thetas, theta_w):
    """Rotation matrices [O,3,9] and weights [O] for the psi x phi x theta
    product, psi outermost (tools/comparison.py:836-841, detector.py:234-244)."""
    psis, phis, thetas = (np.asarray(a, dtype=np.float64) for a in (psis, phis, thetas))
    ang = np.stack(np.meshgrid(psis, phis, thetas, indexing="ij"), axis=-1).reshape(-1, 3)
    rad = np.radians(ang)
    cs = np.empty((ang.shape[0], 6), dtype=np.float64)
    cs[:, 0::2] = np.cos(rad)
    cs[:, 1::2] = np.sin(rad)
    R = np.empty((ang.shape[0], 3, 9), dtype=np.float64)
    corners = np.ascontiguousarray(corners, dtype=np.float64)
    call("gx_host_orientation_matrices", ptr(corners), ptr(cs), int(ang.shape[0]), ptr(R))
    pw, fw, tw = (np.asarray(a, dtype=np.float64) for a in (psi_w, phi_w, theta_w))
    w = ((pw[:, None, None] * fw[None, :, None]) * tw[None, None, :]).reshape(-1)
    return R, np.ascontiguousarray(w)


def grid_corners(det_x, det_y, det_z):
    """p[0,0], p[0,-1], p[-1,0] as rows (detector.py:58-60)."""
    if isinstance(det_x, torch.Tensor):
        g = torch.stack([det_x, det_y, det_z])
        return torch.stack([g[:, 0, 0], g[:, 0, -1], g[:, -1, 0]]).cpu().numpy()
    return np.array([[g[0, 0] for g in (det_x, det_y, det_z)],
                     [g[0, -1] for g in (det_x, det_y, det_z)],
                     [g[-1, 0] for g in (det_x, det_y, det_z)]], dtype=np.float64)


def affine_plan_host(shape, mins, dq, corners, dev3, rows, cols, R, w):
    """Host half of the fixed-point detector kernel (gx_host_affine_orientations): records and
    plan for a [rows, cols] grid with exact corners `corners` ([3,3]: p[0,0], p[0,-1], p[-1,0]) whose
    deviation from their interpolation is at most dev3 per component.  None when unsupported."""
    n = int(len(w))
    Vy, Vx, Vz = (int(v) for v in shape)
    R = np.ascontiguousarray(R, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    rec = np.zeros(n * call("gx_affine_record_bytes"), dtype=np.uint8)
    plan = np.zeros(call("gx_affine_plan_doubles"), dtype=np.float64)
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(9)
    dev3 = np.ascontiguousarray(dev3, dtype=np.float64)
    try:
        call("gx_host_affine_orientations", ptr(corners), ptr(dev3), int(rows), int(cols), ptr(R), ptr(w), n,
             float(mins[0]), float(mins[1]), float(mins[2]), float(dq), Vy, Vx, Vz, ptr(rec), ptr(plan))
    except _lib.GxError as e:
        if e.code != _lib.GX_ERR_UNSUPPORTED:
            raise
        return None
    return corners, rec, plan


_affine_fit_cache = {}
AFFINE_RECORD = np.dtype([("o", "<f8", 3), ("u", "<f8", 3), ("v", "<f8", 3), ("U", "<i4", 3), ("V", "<i4", 3),
                          ("w", "<f4"), ("n_const", "<i4")])
AFFINE_TILE = (16, 32)          # rows, cols of a CTA tile (GA_TH, GA_TW in gx_detector_affine.cu)


class DetectorEngine:
    """Voxel grid resident on the device + accumulation of detector images."""

    def __init__(self, iq, qx, qy, qz, device=None):
        self.device = resolve_device(device)
        with torch.cuda.device(self.device):
            if isinstance(iq, torch.Tensor):
                self.iq = iq.to(self.device, torch.float32).contiguous()
            else:
                self.iq = _dev(np.asarray(iq), self.device, torch.float32).contiguous()
        self.shape = tuple(int(s) for s in self.iq.shape)
        qx, qy, qz = (np.asarray(a, dtype=np.float64) for a in (qx, qy, qz))
        self.mins = (float(np.min(qx)), float(np.min(qy)), float(np.min(qz)))
        self.dq = float(np.diff(qz)[0])               # detector.py:213

    last_slow_fraction = None
    last_kernel = None
    last_plan = None

    @staticmethod
    def _mostly_edge_locked(fast, pmax):
        """True when, for most orientations, some voxel coordinate is (nearly) constant over the
        whole detector and within the fp32 bound of an integer, i.e. every pixel would fall back
        to the exact chain anyway.  Only a dispatch hint: both kernels give identical indices."""
        rec = fast.reshape(-1, call("gx_fast_record_bytes"))[:, :60].copy().view(np.float32)
        m, off, slack = rec[:, :9].reshape(-1, 3, 3), rec[:, 9:12], rec[:, 12:15]
        spread = (np.abs(m) * np.asarray(pmax, dtype=np.float32)[None, None, :]).sum(axis=2)
        near = np.abs(off - np.round(off)) <= (0.5 - slack) + spread
        locked = ((spread < 0.25) & near).any(axis=1)
        return locked.mean() > 0.5

    def affine_plan(self, px, py, pz, R, w):
        """Fixed-point affine model of a [rows, cols] device grid for the orientations R, w:
        (corners [3,3], records uint8 [n * record_bytes], plan float64 [8]) or None when the
        grid is not affine enough / does not fit the fixed-point format."""
        rows, cols = (int(s) for s in px.shape)
        # the fit (one pass over the grid + a device->host read) is cached on the identity and
        # version of the three tensors: drivers reuse one base grid for every call
        key = tuple((g.data_ptr(), g._version) for g in (px, py, pz)) + (rows, cols, str(self.device))
        hit = _affine_fit_cache.get(key)
        if hit is None:
            corners = np.zeros(9, dtype=np.float64)
            dev3 = np.zeros(3, dtype=np.float64)
            scratch = torch.zeros(3, dtype=torch.float64, device=self.device)
            call("gx_grid_affine_fit", ptr(px), ptr(py), ptr(pz), rows, cols, ptr(scratch), ptr(corners),
                 ptr(dev3), _stream())
            while len(_affine_fit_cache) >= 4:
                _affine_fit_cache.pop(next(iter(_affine_fit_cache)))
            hit = _affine_fit_cache[key] = (corners, dev3, (px, py, pz))    # tensors kept alive: pointers stay valid
        return self.affine_plan_host(hit[0], hit[1], rows, cols, R, w)

    def affine_plan_host(self, corners, dev3, rows, cols, R, w):
        return affine_plan_host(self.shape, self.mins, self.dq, corners, dev3, rows, cols, R, w)

    def padded_grid(self):
        """Copy of the voxel grid with rows padded to a multiple of 4 floats (tensor-map strides are
        multiples of 16 bytes), built once per engine for the TMA-brick variant of the gather."""
        if getattr(self, "_padded", None) is None:
            Vy, Vx, Vz = self.shape
            pad = torch.zeros((Vy, Vx, (Vz + 3) // 4 * 4), dtype=torch.float32, device=self.device)
            pad[:, :, :Vz] = self.iq
            self._padded = pad
        return self._padded

    @staticmethod
    def _brick_fits(rec, plan):
        """Every 32 x 16-pixel tile spans fewer than 7 voxels along every axis for every orientation record
        (the 8^3 brick then holds every voxel the tile can touch)."""
        r = np.frombuffer(rec, dtype=AFFINE_RECORD)
        span = (AFFINE_TILE[1] - 1) * np.abs(r["U"].astype(np.float64)) + (AFFINE_TILE[0] - 1) * np.abs(r["V"].astype(np.float64))
        return bool(span.max() / 2.0 ** float(plan[0]) < 7.0)

    def accumulate(self, det_x, det_y, det_z, R, w, image=None, probe=-1, exact_only=None, count_slow=False,
                   kernel=None):
        """image[P,P] (fp64, device) += sum_o w_o * iq[voxel(R_o p)].

        kernel: None (auto) | "affine" | "filtered" | "exact".  Auto uses the
        fixed-point affine-grid kernel for 2-D grids that are affine in (row, col)
        (every grid make_detector + rotations produce), else the fp32-filtered
        generic kernel, else the all-fp64 one.  All three give bit-identical voxel
        indices.  exact_only=True / GIWAXS_B200_EXACT_DETECTOR=1 forces "exact"."""
        dev = self.device
        if exact_only or (exact_only is None and EXACT_DETECTOR_ONLY):
            kernel = "exact"
        with torch.cuda.device(dev):
            shape = tuple(det_x.shape)
            n_pix = int(np.prod(shape))
            on_device = isinstance(det_x, torch.Tensor)
            px, py, pz = ((g.contiguous() if on_device else _dev(np.asarray(g, dtype=np.float64), dev))
                          for g in (det_x, det_y, det_z))
            if image is None:
                image = torch.zeros(n_pix, dtype=torch.float64, device=dev)
            index = torch.empty(n_pix, dtype=torch.int64, device=dev) if probe >= 0 else None
            R = np.ascontiguousarray(R, dtype=np.float64)
            w = np.ascontiguousarray(w, dtype=np.float64)
            d_R = _dev(R, dev)
            Vy, Vx, Vz = self.shape
            slow = torch.zeros(1, dtype=torch.int64, device=dev) if count_slow else None
            self.last_slow_fraction = None

            if kernel in (None, "affine", "affine_tma") and len(shape) == 2 and shape[0] > 1 and shape[1] > 1:
                # The host model of a chunk of orientations is built while the GPU works on the
                # previous chunk (launches are asynchronous): a short first chunk gets the device
                # busy at once, later chunks are sized so that their host time stays hidden.
                n = len(w)
                bounds = [0, min(n, 48)]
                while bounds[-1] < n:
                    bounds.append(min(n, bounds[-1] + 1024))
                launched = False
                for b0, b1 in zip(bounds[:-1], bounds[1:]):
                    Rc, wc = R[b0:b1], w[b0:b1]
                    plan = self.affine_plan(px, py, pz, Rc, wc)
                    # edge-locked: a coordinate that is constant over the detector, sits on a voxel edge
                    # and could not be modelled -> every pixel would take the exact path anyway
                    ok = plan is not None and (kernel in ("affine", "affine_tma") or plan[2][6] <= 0.5 * (b1 - b0))
                    if not ok:
                        if kernel in ("affine", "affine_tma"):
                            raise _lib.GxError(_lib.GX_ERR_UNSUPPORTED, "detector grid is not affine in (row, col)")
                        if not launched:
                            break                          # generic kernels below take the whole set
                        d_w = _dev(wc, dev)                # this chunk only: all-fp64 kernel
                        pr = probe - b0 if b0 <= probe < b1 else -1
                        call("gx_detector_accumulate", ptr(self.iq), Vy, Vx, Vz, self.mins[0], self.mins[1],
                             self.mins[2], self.dq, ptr(px), ptr(py), ptr(pz), n_pix, ptr(d_R[b0:b1]), ptr(d_w),
                             b1 - b0, ptr(image), int(pr), ptr(index) if pr >= 0 else None, _stream())
                        continue
                    corners, rec, pl = plan
                    d_rec = _dev(rec, dev)
                    pr = probe - b0 if b0 <= probe < b1 else -1
                    if (DETECTOR_TMA or kernel == "affine_tma") and pr < 0 and not count_slow and self._brick_fits(rec, pl):
                        pad = self.padded_grid()
                        call("gx_detector_accumulate_affine_brick", ptr(self.iq), ptr(pad), int(pad.shape[2]), Vy, Vx, Vz,
                             self.mins[0], self.mins[1], self.mins[2], self.dq, ptr(px), ptr(py), ptr(pz), shape[0],
                             shape[1], ptr(corners), ptr(d_rec), ptr(d_R[b0:b1]), b1 - b0, ptr(pl), ptr(image), _stream())
                        launched = True
                        self.last_kernel, self.last_plan = "affine_tma", pl
                        continue
                    call("gx_detector_accumulate_affine", ptr(self.iq), Vy, Vx, Vz, self.mins[0], self.mins[1],
                         self.mins[2], self.dq, ptr(px), ptr(py), ptr(pz), shape[0], shape[1], ptr(corners),
                         ptr(d_rec), ptr(d_R[b0:b1]), b1 - b0, ptr(pl), ptr(image), int(pr),
                         ptr(index) if pr >= 0 else None, ptr(slow), _stream())
                    launched = True
                    self.last_kernel, self.last_plan = "affine", pl
                if launched:
                    if count_slow:
                        self.last_slow_fraction = float(slow.item()) / (n_pix * len(w))
                    return image, index

            if kernel in (None, "filtered"):
                if on_device and len(shape) == 2:
                    # driver path: the grid is affine in (row, col), so |component| peaks at a corner
                    c = grid_corners(px, py, pz)
                    corners = np.vstack([c, c[1] + c[2] - c[0]])
                    pmax = np.abs(corners).max(axis=0) * (1.0 + 1e-9)
                else:
                    pmax = np.array([float(g.abs().max()) for g in (px, py, pz)])
                fast = np.zeros(len(w) * call("gx_fast_record_bytes"), dtype=np.uint8)
                pmax = np.ascontiguousarray(pmax, dtype=np.float64)      # keep alive across the call
                call("gx_host_fast_orientations", ptr(R), ptr(w), int(len(w)), self.mins[0], self.mins[1],
                     self.mins[2], self.dq, ptr(pmax), ptr(fast))
                if kernel == "filtered" or not self._mostly_edge_locked(fast, pmax):
                    d_fast = _dev(fast, dev)
                    call("gx_detector_accumulate_fast", ptr(self.iq), Vy, Vx, Vz, self.mins[0], self.mins[1],
                         self.mins[2], self.dq, ptr(px), ptr(py), ptr(pz), n_pix, ptr(d_fast), ptr(d_R),
                         int(len(w)), ptr(image), int(probe), ptr(index), ptr(slow), _stream())
                    torch.cuda.current_stream().synchronize()
                    self.last_kernel = "filtered"
                    if count_slow:
                        self.last_slow_fraction = float(slow.item()) / (n_pix * len(w))
                    return image, index

            # all-fp64 kernel (identity rotation steps skipped)
            d_w = _dev(w, dev)
            call("gx_detector_accumulate", ptr(self.iq), Vy, Vx, Vz, self.mins[0], self.mins[1], self.mins[2],
                 self.dq, ptr(px), ptr(py), ptr(pz), n_pix, ptr(d_R), ptr(d_w), int(len(w)),
                 ptr(image), int(probe), ptr(index), _stream())
            torch.cuda.current_stream().synchronize()
            self.last_kernel = "exact"
        return image, index


def rotate_points(R, gx, gy, gz, device=None):
    """R @ [x;y;z] on the device with the reference's fma chain.  NumPy grids in
    -> NumPy grids out; device tensors in -> device tensors out."""
    dev = resolve_device(device)
    on_device = isinstance(gx, torch.Tensor)
    shape = tuple(gx.shape)
    with torch.cuda.device(dev):
        if on_device:
            src = [g.contiguous().view(-1) for g in (gx, gy, gz)]
        else:
            src = [_dev(np.asarray(g, dtype=np.float64).ravel(), dev) for g in (gx, gy, gz)]
        out = [torch.empty_like(s) for s in src]
        Rh = np.ascontiguousarray(R, dtype=np.float64)
        call("gx_rotate_points", ptr(Rh), ptr(src[0]), ptr(src[1]), ptr(src[2]), int(src[0].numel()),
             ptr(out[0]), ptr(out[1]), ptr(out[2]), _stream())
        if on_device:
            return tuple(o.view(shape) for o in out)
        return tuple(o.cpu().numpy().reshape(shape) for o in out)


def detector_epilogue(image, rows, cols, mirror, device, finish=True):
    with torch.cuda.device(device):
        out = torch.empty(rows * cols, dtype=torch.float64, device=device)
        call("gx_detector_epilogue", ptr(image), int(rows), int(cols), int(bool(mirror)), int(finish),
             ptr(out), _stream())
        torch.cuda.current_stream().synchronize()
    return out.view(rows, cols)
