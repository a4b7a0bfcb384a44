"""CPU restatement of the reference's hot path (torch-CPU / numpy / scipy).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Written from the behaviour of
``/root/reference/src/torch_motion_correction`` (cited per function as ``file:line``);
third-party numerics come from ``oracle.deps``.  Pinned against the unmodified reference
run over the same dependency layer: ``tests/golden/`` + ``tests/test_oracle_golden.py``.

All reference quirks that change results are reproduced on purpose:
Q1 (cached patches mutated by the mask), Q2 (``correct_motion_fast`` negates the caller's
field in place and treats Angstrom as px), Q3, Q4, Q5-Q9, Q10/Q11 (shuffled mini-batches
with per-batch mean weighting), Q13, Q17 (two-stage shift field), Q18.
"""

from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn.functional as F

from oracle import deps

# --------------------------------------------------------------------------------------
# normalisation + filter tables   (utils.py)
# --------------------------------------------------------------------------------------


def normalize_image(image: torch.Tensor, frac_low: float = 0.25, frac_high: float = 0.75):
    """utils.py:49-84 -- scalar mean / unbiased std of the central box of the whole stack."""
    _, h, w = image.shape
    box = image[:, int(frac_low * h) : int(frac_high * h), int(frac_low * w) : int(frac_high * w)]
    std, mean = torch.std_mean(box, dim=(-3, -2, -1))
    return (image - mean) / std


def band_edges(frequency_range, pixel_spacing: float):
    """utils.py:99-102 -- (low, high) cut-offs in cycles/px as fp32 scalars."""
    cuton, cutoff_max = torch.as_tensor(frequency_range).float()
    cutoff = torch.lerp(cuton, cutoff_max, 1.0)
    low = torch.as_tensor(1 / cuton, dtype=torch.float32) * pixel_spacing
    high = torch.as_tensor(1 / cutoff, dtype=torch.float32) * pixel_spacing
    return low, high


def prepare_bandpass_filter(frequency_range, patch_shape, pixel_spacing: float):
    """utils.py:87-114 -- hard band ``low < f <= high`` (falloff 0) on the rfft grid."""
    low, high = band_edges(frequency_range, pixel_spacing)
    return deps.bandpass_filter(low, high, 0, tuple(patch_shape), rfft=True, fftshift=False)


def fourier_weight(shape, pixel_spacing: float, b_factor: float, frequency_range):
    """bandpass * b_envelope as applied at estimate_motion_xc.py:341,346 / optimizer:504-508."""
    band = prepare_bandpass_filter(frequency_range, shape, pixel_spacing)
    env = deps.b_envelope(b_factor, tuple(shape), pixel_spacing, rfft=True, fftshift=False)
    return band, env


# --------------------------------------------------------------------------------------
# patch geometry   (patch_grid/_patch_grid_centers.py, _patch_grid_indices.py)
# --------------------------------------------------------------------------------------


def patch_centers_1d(dim_length: int, patch_length: int, patch_step: int, distribute: bool = True):
    """_patch_grid_centers.py:70-111."""
    lo = patch_length // 2
    hi = max(dim_length - lo - 1, lo)
    centers = torch.arange(lo, hi + 1, patch_step)
    if distribute:
        gap = hi - centers[-1]
        centers = centers + torch.round(torch.linspace(0, gap, len(centers))).long()
    return centers


def patch_grid_centers(image_shape, patch_shape, patch_step, distribute: bool = True):
    """_patch_grid_centers.py:10-67,171-213 -- (t, gh, gw, 3) int64 (t, y, x) centres."""
    cd, ch, cw = (
        patch_centers_1d(n, p, s, distribute) for n, p, s in zip(image_shape, patch_shape, patch_step)
    )
    grid = torch.stack(torch.meshgrid(cd, ch, cw, indexing="ij"), dim=-1)
    return grid


def extract_frame_patches(frame: torch.Tensor, cy: torch.Tensor, cx: torch.Tensor, p: int):
    """_patch_grid_indices.py:75-97 + _patch_grid.py:464-477 -- (gh, gw, p, p) copy."""
    off = torch.arange(p) - p // 2
    iy = (cy[:, None] + off[None, :])[:, None, :, None]
    ix = (cx[:, None] + off[None, :])[None, :, None, :]
    return frame[iy, ix]


class LazyCacheModel:
    """Bookkeeping model of ``LazyPatchGrid``'s cache (``_patch_grid.py:264-347``).

    Only tracks *how many times the mask has been multiplied into* each cached frame
    (quirk Q1): ``get(j)`` returns that count and (re)creates an entry with count 0 on a
    miss, evicting half the keys when more than 50 are held, exactly as the reference's
    ``dict`` + ``set`` pair does (same container types => same iteration order).
    """

    LIMIT = 50

    def __init__(self):
        self.count: dict[int, int] = {}
        self.keys: set[int] = set()

    def touch(self, j: int) -> int:
        if j in self.count:
            return self.count[j]
        self.count[j] = 0
        self.keys.add(j)
        if len(self.count) > self.LIMIT:
            victims = list(self.keys)[: len(self.keys) // 2]
            for v in victims:
                self.count.pop(v, None)
                self.keys.discard(v)
            # NB: the reference returns the freshly extracted tensor even if its own key
            # was just evicted; the returned object is then un-cached (mutations are lost).
        return 0

    def mask_in_place(self, j: int) -> None:
        if j in self.count:
            self.count[j] += 1


def q1_schedule(t: int, strategy: str, reference_frame: int):
    """Mask-exponent schedule implied by Q1 for every processed frame.

    Returns ``(ref_exponents, order)``: for ``mean_except_current`` ``ref_exponents[k]`` is a
    length-t int list (exponent of the *inner* mask on frame j inside frame k's reference,
    entry k unused); for ``middle_frame`` it is a single int n such that the reference is
    ``mask**n * P_ref``.  estimate_motion_xc.py:297-346.
    """
    cache = LazyCacheModel()
    sched = {}
    for k in range(t):
        if strategy == "middle_frame":
            if k == reference_frame:
                continue
            e = cache.touch(reference_frame)
            cache.mask_in_place(reference_frame)  # ``ref_patches *= mask`` on the cached view
            # if the entry was evicted at creation the mutation hits an un-cached tensor
            sched[k] = e + 1
            cache.touch(k)
            cache.mask_in_place(k)
        elif strategy == "mean_except_current":
            row = [0] * t
            for j in range(t):
                if j != k:
                    row[j] = cache.touch(j)
            sched[k] = row
            cache.touch(k)
            cache.mask_in_place(k)
        else:
            raise ValueError(f"Unknown reference_strategy: {strategy}")
    return sched


# --------------------------------------------------------------------------------------
# spline field utilities   (deformation_field_utils.py)
# --------------------------------------------------------------------------------------


def _matrix(grid_type: str):
    if grid_type == "catmull_rom":
        return deps.CATMULL_ROM_MATRIX
    if grid_type == "bspline":
        return deps.BSPLINE_MATRIX
    raise ValueError(f"Invalid grid type: {grid_type}")


def evaluate_deformation_field(field: torch.Tensor, tyx: torch.Tensor, grid_type: str = "catmull_rom"):
    """deformation_field_utils.py:9-39."""
    return deps.evaluate_cubic_grid_3d(field.to(torch.float32), tyx, _matrix(grid_type))


def evaluate_deformation_field_at_t(field, t: float, grid_shape, grid_type: str = "catmull_rom"):
    """deformation_field_utils.py:42-93 -- (2, h, w) lattice of shifts at one time."""
    h, w = grid_shape
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    tyx = torch.stack([torch.full_like(yy, float(t)), yy, xx], dim=-1).reshape(-1, 3)
    out = evaluate_deformation_field(field, tyx, grid_type)
    return out.reshape(h, w, -1).permute(2, 0, 1)


def resample_deformation_field(field: torch.Tensor, target_resolution):
    """deformation_field_utils.py:96-126 -- always Catmull-Rom (Q3)."""
    nt, nh, nw = target_resolution
    tt, yy, xx = torch.meshgrid(
        torch.linspace(0, 1, nt), torch.linspace(0, 1, nh), torch.linspace(0, 1, nw), indexing="ij"
    )
    tyx = torch.stack([tt, yy, xx], dim=-1)
    return evaluate_deformation_field(field, tyx).permute(3, 0, 1, 2)


def image_shifts_to_deformation_field(shifts: torch.Tensor, pixel_spacing: float):
    """deformation_field_utils.py:129-162 -- (t, 2) px -> (2, t, 1, 1) Angstrom."""
    return (shifts * pixel_spacing).transpose(0, 1)[:, :, None, None]


# --------------------------------------------------------------------------------------
# correction   (correct_motion.py)
# --------------------------------------------------------------------------------------


def get_pixel_shifts(frame, pixel_spacing: float, lattice: torch.Tensor, pixel_grid: torch.Tensor):
    """correct_motion.py:132-185 -- bicubic/reflection lookup of the (2, gh, gw) lattice."""
    h, w = frame.shape
    _, gh, gw = lattice.shape
    img_len = torch.as_tensor([h - 1, w - 1], dtype=torch.float32)
    lat_len = torch.as_tensor([gh - 1, gw - 1], dtype=torch.float32)
    q = (pixel_grid / img_len) * lat_len
    grid = deps.array_to_grid_sample(q, (gh, gw))
    s = F.grid_sample(lattice[None], grid[None], mode="bicubic", padding_mode="reflection", align_corners=True)
    return (s / pixel_spacing)[0].permute(1, 2, 0)


def correct_frame(frame, pixel_spacing: float, lattice: torch.Tensor):
    """correct_motion.py:81-129."""
    pixel_grid = deps.coordinate_grid(frame.shape)
    shifts = get_pixel_shifts(frame, pixel_spacing, lattice, pixel_grid)
    return deps.sample_image_2d(frame, pixel_grid + shifts, interpolation="bicubic")


def correct_motion(image, deformation_grid, pixel_spacing: float, grid_type: str = "catmull_rom"):
    """correct_motion.py:18-78 -- (t, h, w) warped stack (10x lattice then bicubic, Q17)."""
    t = image.shape[0]
    gh, gw = deformation_grid.shape[-2:]
    times = torch.linspace(0, 1, steps=t)
    out = [
        correct_frame(
            frame, pixel_spacing, evaluate_deformation_field_at_t(deformation_grid, ft, (10 * gh, 10 * gw), grid_type)
        )
        for frame, ft in zip(image, times)
    ]
    return torch.stack(out, dim=0)


def correct_motion_two_grids(image, new_data, base_data, pixel_spacing: float, grid_type: str = "catmull_rom"):
    """correct_motion.py:188-317 (forward value): lattice = new(tyx) + base(tyx)."""
    t = image.shape[0]
    gh, gw = new_data.shape[-2:]
    times = torch.linspace(0, 1, steps=t)
    out = []
    for frame, ft in zip(image, times):
        lat = evaluate_deformation_field_at_t(new_data, ft, (10 * gh, 10 * gw), grid_type) + evaluate_deformation_field_at_t(
            base_data, ft, (10 * gh, 10 * gw), grid_type
        )
        out.append(correct_frame(frame, pixel_spacing, lat))
    return torch.stack(out, dim=0)


def correct_motion_slow(image, deformation_grid):
    """correct_motion.py:320-427 -- exact per-pixel Catmull-Rom spline, no /pixel_spacing."""
    t, h, w = image.shape
    times = torch.linspace(0, 1, steps=t)
    pixel_grid = deps.coordinate_grid((h, w))
    norm = pixel_grid / torch.as_tensor([h - 1, w - 1], dtype=torch.float32)
    out = []
    for frame, ft in zip(image, times):
        tyx = F.pad(norm, (1, 0), value=float(ft))
        s = evaluate_deformation_field(deformation_grid, tyx)
        out.append(deps.sample_image_2d(frame, pixel_grid + s, interpolation="bicubic"))
    return torch.stack(out, dim=0)


def correct_motion_fast(image, deformation_grid):
    """correct_motion.py:430-498 -- rigid Fourier shift; NEGATES ``deformation_grid`` in place (Q2)."""
    if tuple(deformation_grid.shape[-2:]) != (1, 1):
        raise ValueError(
            f"Expected single patch deformation field with shape (2, t, 1, 1), but got shape {deformation_grid.shape}."
        )
    t, h, w = image.shape
    deformation_grid *= -1  # Q2: the reference's ``shifts`` is a view of the caller's tensor
    shifts = deformation_grid[:, :, 0, 0].transpose(0, 1)
    dft = torch.fft.rfftn(image, dim=(-2, -1))
    dft = deps.fourier_shift_dft_2d(dft, (h, w), shifts, rfft=True, fftshifted=False)
    return torch.fft.irfftn(dft, s=(h, w))


# --------------------------------------------------------------------------------------
# cross-correlation estimators   (estimate_motion_xc.py)
# --------------------------------------------------------------------------------------


def dose_weight(movie: torch.Tensor, pixel_size: float, pre_exposure: float = 0.0, dose_per_frame: float = 1.0,
                voltage: float = 300.0):
    """Dose-weighted sum of an aligned stack, as the reference's example script does it
    (``examples/ttMotion.py:331-351``): rfft2 (ortho) -> ``dose_weight_movie`` -> irfft2 (ortho) -> sum over frames.
    PARITY UNPINNED: ``torch_fourier_filter.dose_weight`` is restated in ``oracle/deps.py``."""
    shape = (movie.shape[-2], movie.shape[-1])
    dft = torch.fft.rfft2(movie, dim=(-2, -1), norm="ortho")
    dw = deps.dose_weight_movie(dft, shape, pixel_size, pre_exposure, dose_per_frame, voltage, -1, True, False)
    return torch.fft.irfft2(dw, s=shape, dim=(-2, -1), norm="ortho").sum(dim=0)


def _wrap(peak, n: int):
    return torch.where(peak <= n // 2, peak, peak - n)


def estimate_global_motion(image, pixel_spacing: float, reference_frame=None, b_factor: float = 500, frequency_range=(300, 10)):
    """estimate_motion_xc.py:21-135 -- integer-pixel whole-frame XC vs one frame (Q5)."""
    t, h, w = image.shape
    if reference_frame is None:
        reference_frame = t // 2
    image = normalize_image(image)
    mask = deps.circle(min(h, w) / 4, (h, w), smoothing_radius=min(h, w) / 8)
    band, env = fourier_weight((h, w), pixel_spacing, b_factor, frequency_range)
    spec = torch.fft.rfftn(image * mask, dim=(-2, -1)) * band * env
    ref = spec[reference_frame]
    shifts = torch.zeros((t, 2))
    for k in range(t):
        if k == reference_frame:
            continue
        cc = torch.fft.irfftn(torch.conj(ref) * spec[k], s=(h, w))
        py, px = divmod(int(torch.argmax(cc.flatten())), w)
        shifts[k, 0] = py if py <= h // 2 else py - h
        shifts[k, 1] = px if px <= w // 2 else px - w
    return image_shifts_to_deformation_field(shifts, pixel_spacing)


def sub_pixel_refinement(cc: torch.Tensor, peak_idx: torch.Tensor):
    """estimate_motion_xc.py:414-483 -- 3-point parabola per axis (Q6). cc (n, p, p)."""
    n, ph, pw = cc.shape
    py = (peak_idx // pw).float()
    px = (peak_idx % pw).float()
    for i in range(n):
        y, x = int(peak_idx[i]) // pw, int(peak_idx[i]) % pw
        if 1 <= y < ph - 1 and 1 <= x < pw - 1:
            a, b, c = cc[i, y - 1, x], cc[i, y, x], cc[i, y + 1, x]
            if c != a:
                py[i] += 0.5 * (a - c) / (a - 2 * b + c)
            a, b, c = cc[i, y, x - 1], cc[i, y, x], cc[i, y, x + 1]
            if c != a:
                px[i] += 0.5 * (a - c) / (a - 2 * b + c)
    return py, px


def outlier_rejection(sy: torch.Tensor, sx: torch.Tensor, threshold: float):
    """estimate_motion_xc.py:538-627 -- lower median / unbiased std z-score (Q8)."""
    shape = sy.shape
    sy, sx = sy.flatten(), sx.flatten()
    tiny = torch.tensor(1e-6)
    zy = (sy - torch.median(sy)).abs() / torch.max(torch.std(sy), tiny)
    zx = (sx - torch.median(sx)).abs() / torch.max(torch.std(sx), tiny)
    bad = (zy > threshold) | (zx > threshold)
    good = ~bad
    my = sy[good].mean() if good.any() else torch.median(sy)
    mx = sx[good].mean() if good.any() else torch.median(sx)
    sy, sx = sy.clone(), sx.clone()
    sy[bad] = my
    sx[bad] = mx
    return sy.view(shape), sx.view(shape)


def temporal_smoothing(field: torch.Tensor, window: int):
    """estimate_motion_xc.py:486-535 -- Savitzky-Golay, polyorder 1, scipy mode='interp' (Q9)."""
    from scipy.signal import savgol_filter

    if window % 2 == 0:
        window += 1
    window = min(window, field.shape[1])
    if window < 3:
        return field
    out = field.clone()
    for gy in range(field.shape[2]):
        for gx in range(field.shape[3]):
            for c in range(2):
                series = field[c, :, gy, gx].cpu().numpy()
                out[c, :, gy, gx] = torch.from_numpy(savgol_filter(series, window, 1)).to(field.device)
    return out


def estimate_motion_cross_correlation_patches(
    image,
    pixel_spacing: float,
    reference_frame=None,
    reference_strategy: str = "mean_except_current",
    b_factor: float = 500,
    frequency_range=(300, 10),
    patch_sidelength: int = 1024,
    sub_pixel: bool = True,
    smooth: bool = True,
    smoothing_window_size: int = 5,
    deformation_field=None,
    reject_outliers: bool = True,
    outlier_threshold: float = 3.0,
    return_raw: bool = False,
    dose=None,
):
    """estimate_motion_xc.py:138-411 incl. Q1 (aliased cache), Q2/Q3 pre-correction, Q4.

    ``dose = (pre_exposure, dose_per_frame, voltage)`` (NOT in the reference; additive option of the B200 build,
    parity unpinned): every frame's masked patch spectra are multiplied by ``dose_weight_movie``'s exposure filter
    before the leave-one-out mean (formed in Fourier space, which is the same linear operation) and the product."""
    t, h, w = image.shape
    if reference_frame is None:
        reference_frame = t // 2
    image = normalize_image(image)
    if deformation_field is not None:
        if tuple(deformation_field.shape[-2:]) == (1, 1):
            image = correct_motion_fast(image, deformation_field)  # mutates the caller's field (Q2)
        else:
            image = correct_motion(image, deformation_field, pixel_spacing, grid_type="bspline")
    p = patch_sidelength
    centers = patch_grid_centers((t, h, w), (1, p, p), (1, p // 2, p // 2))
    gh, gw = centers.shape[1:3]
    cy, cx = centers[0, :, 0, 1], centers[0, 0, :, 2]
    mask = deps.circle(p / 4, (p, p), smoothing_radius=p / 8)
    band, env = fourier_weight((p, p), pixel_spacing, b_factor, frequency_range)
    if deformation_field is None:
        field = torch.zeros((2, t, gh, gw))
    else:
        field = resample_deformation_field(deformation_field, (t, gh, gw)).clone()

    # the reference's cache holds one (gh, gw, p, p) tensor per frame, mutated in place by
    # ``frame_patches *= mask`` (and by ``ref_patches *= mask`` for ``middle_frame``).
    # Reproduce by keeping "current state" tensors and following the cache model.
    cache = LazyCacheModel()
    state: dict[int, torch.Tensor] = {}

    def fetch(j: int) -> torch.Tensor:
        was_cached = j in cache.count
        cache.touch(j)
        if not was_cached:
            state[j] = extract_frame_patches(image[j], cy, cx, p)
            for dead in [q for q in state if q not in cache.count and q != j]:
                del state[dead]
        return state[j]

    raw = torch.zeros((2, t, gh, gw))
    q = None
    if dose is not None:
        ones = torch.ones((t, p, p // 2 + 1), dtype=torch.complex64)
        q = deps.dose_weight_movie(ones, (p, p), pixel_spacing, dose[0], dose[1], dose[2], -1, True, False).real
    for k in range(t):
        rf_dose = None
        if reference_strategy == "middle_frame":
            if k == reference_frame:
                continue
            ref = fetch(reference_frame)
            ref *= mask  # in place on the cached tensor (Q1)
            ref_m = ref
            if q is not None:
                rf_dose = torch.fft.rfftn(ref_m, dim=(-2, -1)) * q[reference_frame]
        elif reference_strategy == "mean_except_current" and q is not None:
            acc_f, count = None, 0
            for j in range(t):
                if j == k:
                    continue
                term = torch.fft.rfftn(fetch(j) * mask, dim=(-2, -1)) * q[j]
                acc_f = term if acc_f is None else acc_f + term
                count += 1
            rf_dose = acc_f / count
            ref_m = None
        elif reference_strategy == "mean_except_current":
            acc, count = None, 0
            for j in range(t):
                if j == k:
                    continue
                other = fetch(j)
                acc = other.clone() if acc is None else acc.add_(other)
                count += 1
            ref_m = (acc / count) * mask
        else:
            raise ValueError(f"Unknown reference_strategy: {reference_strategy}")
        cur = fetch(k)
        cur *= mask  # Q1: mutates the cache entry
        if rf_dose is not None:
            rf = rf_dose * band * env
            ff = torch.fft.rfftn(cur, dim=(-2, -1)) * band * env * q[k]
        else:
            rf = torch.fft.rfftn(ref_m, dim=(-2, -1)) * band * env
            ff = torch.fft.rfftn(cur, dim=(-2, -1)) * band * env
        cc = torch.fft.irfftn(torch.conj(rf) * ff, s=(p, p)).reshape(gh * gw, p, p)
        peak = torch.argmax(cc.reshape(gh * gw, -1), dim=1)
        if sub_pixel:
            py, px = sub_pixel_refinement(cc, peak)
        else:
            py, px = peak // p, peak % p
        sy, sx = _wrap(py, p).view(gh, gw), _wrap(px, p).view(gh, gw)
        raw[0, k], raw[1, k] = sy, sx
        if reject_outliers:
            sy, sx = outlier_rejection(sy, sx, outlier_threshold)
        field[0, k] += sy * pixel_spacing
        field[1, k] += sx * pixel_spacing
    if smooth:
        field = temporal_smoothing(field, smoothing_window_size)
    field = field - torch.mean(field)  # Q4: one joint scalar mean
    if return_raw:
        return field, centers, raw
    return field, centers


# --------------------------------------------------------------------------------------
# spline-coefficient optimiser   (estimate_motion_optimizer.py, patch_utils.py)
# --------------------------------------------------------------------------------------


def patch_batches(centers: torch.Tensor, image_shape, batch_size: int, randomized: bool = True):
    """patch_utils.py:147-190 -- shuffled (random.shuffle) mini-batches of patch indices and
    their (t, b, 3) normalised centres (Q10: time-major despite the docs)."""
    t, gh, gw, _ = centers.shape
    T, H, W = image_shape
    norm = centers.clone().float()
    norm[..., 0] /= float(T - 1)
    norm[..., 1] /= float(H - 1)
    norm[..., 2] /= float(W - 1)
    norm = norm.reshape(t, -1, 3)
    order = list(range(gh * gw))
    if randomized:
        random.shuffle(order)
    for i in range(0, gh * gw, batch_size):
        sel = order[i : i + batch_size]
        yield sel, norm[:, sel]


def shifted_patch_spectra(new_data, base_data, spectra, centers_norm, pixel_spacing, p, band, env, grid_type):
    """estimate_motion_optimizer.py:442-510 -- S = F exp(-2 pi i f.s) band env, s = -(new+base)(c)/px."""
    m = _matrix(grid_type)
    pred = -1 * (deps.evaluate_cubic_grid_3d(new_data, centers_norm, m) + deps.evaluate_cubic_grid_3d(base_data, centers_norm, m))
    pred = pred.transpose(0, 1)  # the reference's mislabelled rearrange: (t, b, 2) -> (b, t, 2)
    s = deps.fourier_shift_dft_2d(spectra, (p[0], p[1]), pred / pixel_spacing, rfft=True, fftshifted=False)
    return s * band * env


def batch_loss(shifted: torch.Tensor, ph: int, pw: int, loss_type: str = "mse"):
    """estimate_motion_optimizer.py:391-404,611-671 -- leave-one-out reference + loss (Q13)."""
    t = shifted.shape[1]
    total = shifted.sum(dim=1, keepdim=True)
    ref = (total - shifted) / (t - 1) if t > 1 else shifted
    if loss_type == "mse":
        return torch.mean((shifted - ref).abs() ** 2) / (ph * pw)
    x = torch.fft.irfftn(shifted, s=(ph, pw), dim=(-2, -1))
    y = torch.fft.irfftn(ref, s=(ph, pw), dim=(-2, -1))
    if loss_type == "cc":
        return -(x * y).sum(dim=(-2, -1)).mean()
    if loss_type == "ncc":
        xc = x - x.mean(dim=(-2, -1), keepdim=True)
        yc = y - y.mean(dim=(-2, -1), keepdim=True)
        den = torch.sqrt((xc.square().sum(dim=(-2, -1)) + 1e-8) * (yc.square().sum(dim=(-2, -1)) + 1e-8))
        return -((xc * yc).sum(dim=(-2, -1)) / den).mean()
    raise ValueError(loss_type)


def local_motion_problem(image, pixel_spacing, patch_shape, resolution, initial_field, b_factor, frequency_range, grid_type):
    """Set-up shared by ``estimate_local_motion`` and the loss/gradient probes:
    estimate_motion_optimizer.py:96-191."""
    t, h, w = image.shape
    ph, pw = patch_shape
    image = normalize_image(image)
    centers = patch_grid_centers((t, h, w), (1, ph, pw), (1, ph // 2, pw // 2))
    if initial_field is None:
        base = torch.zeros((2, *resolution))
    else:
        base = resample_deformation_field(initial_field.detach(), tuple(resolution)).clone()
        base -= torch.mean(base)
    mask = deps.circle(pw / 4, (ph, pw), smoothing_radius=pw / 4)  # Q18
    band, env = fourier_weight((ph, pw), pixel_spacing, b_factor, frequency_range)
    return image, centers, base, mask, band, env


def loss_and_grad(
    image, pixel_spacing, patch_shape, resolution, initial_field, new_data, batches,
    b_factor=500, frequency_range=(300, 10), grid_type="catmull_rom", loss_type="mse",
):
    """One full pass of estimate_motion_optimizer.py:362-412 for a GIVEN list of mini-batches
    (lists of flat patch indices): returns (sum of batch losses, d/d new_data)."""
    image, centers, base, mask, band, env = local_motion_problem(
        image, pixel_spacing, patch_shape, resolution, initial_field, b_factor, frequency_range, grid_type
    )
    t, h, w = image.shape
    ph, pw = patch_shape
    norm = centers.clone().float()
    norm[..., 0] /= float(t - 1)
    norm[..., 1] /= float(h - 1)
    norm[..., 2] /= float(w - 1)
    norm = norm.reshape(t, -1, 3)
    flat = centers[0].reshape(-1, 3)
    new = new_data.clone().detach().requires_grad_(True)
    total = 0.0
    for sel in batches:
        patches = torch.stack(
            [image[:, int(flat[i, 1]) - ph // 2 : int(flat[i, 1]) - ph // 2 + ph, int(flat[i, 2]) - pw // 2 : int(flat[i, 2]) - pw // 2 + pw] for i in sel]
        )
        spectra = torch.fft.rfftn(patches * mask, dim=(-2, -1))
        s = shifted_patch_spectra(new, base, spectra, norm[:, sel], pixel_spacing, (ph, pw), band, env, grid_type)
        loss = batch_loss(s, ph, pw, loss_type)
        loss.backward()
        total += float(loss.detach())
    return total, new.grad.detach().clone()


def estimate_local_motion(
    image, pixel_spacing, patch_shape, deformation_field_resolution, initial_deformation_field,
    n_iterations: int = 100, b_factor: float = 500, frequency_range=(300, 10), optimizer_type: str = "adam",
    grid_type: str = "catmull_rom", loss_type: str = "mse", optimizer_kwargs=None, return_losses: bool = False,
):
    """estimate_motion_optimizer.py:28-439 (non-LBFGS branch :361-429 and LBFGS :219-359)."""
    kw = dict(optimizer_kwargs or {})
    image, centers, base, mask, band, env = local_motion_problem(
        image, pixel_spacing, patch_shape, deformation_field_resolution, initial_deformation_field,
        b_factor, frequency_range, grid_type,
    )
    t, h, w = image.shape
    ph, pw = patch_shape
    flat = centers[0].reshape(-1, 3)
    new = torch.zeros((2, *deformation_field_resolution), requires_grad=True)
    opt = make_optimizer(optimizer_type, [new], kw)

    def spectra_of(sel):
        patches = torch.stack(
            [image[:, int(flat[i, 1]) - ph // 2 : int(flat[i, 1]) - ph // 2 + ph, int(flat[i, 2]) - pw // 2 : int(flat[i, 2]) - pw // 2 + pw] for i in sel]
        )
        return torch.fft.rfftn(patches * mask, dim=(-2, -1))

    losses = []
    for _ in range(n_iterations):
        if optimizer_type.lower() == "lbfgs":
            sub = kw.get("lbfgs_patch_subsample", None)

            def closure():
                opt.zero_grad()
                acc, n = None, 0
                for idx, (sel, cn) in enumerate(patch_batches(centers, (t, h, w), 1, True)):
                    if sub is not None and idx >= sub:
                        break
                    s = shifted_patch_spectra(new, base, spectra_of(sel), cn, pixel_spacing, (ph, pw), band, env, grid_type)
                    l = batch_loss(s, ph, pw, loss_type)
                    acc = l if acc is None else acc + l
                    n += 1
                avg = acc / n
                avg.backward()
                return avg

            losses.append(float(opt.step(closure).detach()))
        else:
            total, n = 0.0, 0
            for sel, cn in patch_batches(centers, (t, h, w), 8, True):
                s = shifted_patch_spectra(new, base, spectra_of(sel), cn, pixel_spacing, (ph, pw), band, env, grid_type)
                l = batch_loss(s, ph, pw, loss_type)
                l.backward()
                total += float(l.detach())
                n += 1
            opt.step()
            opt.zero_grad()
            losses.append(total / n)
    final = new.detach() + base
    final = final - torch.mean(final)
    if return_losses:
        return final, losses
    return final


def make_optimizer(optimizer_type: str, params, kw):
    """estimate_motion_optimizer.py:513-608 -- defaults of the four supported optimisers."""
    name = optimizer_type.lower()
    if name == "adam":
        return torch.optim.Adam(params, lr=kw.get("lr", 0.01), betas=kw.get("betas", (0.9, 0.999)), eps=kw.get("eps", 1e-8),
                                weight_decay=kw.get("weight_decay", 0), amsgrad=kw.get("amsgrad", False))
    if name == "sgd":
        return torch.optim.SGD(params, lr=kw.get("lr", 0.01), momentum=kw.get("momentum", 0.9), weight_decay=kw.get("weight_decay", 0),
                               dampening=kw.get("dampening", 0), nesterov=kw.get("nesterov", True))
    if name == "rmsprop":
        return torch.optim.RMSprop(params, lr=kw.get("lr", 0.01), alpha=kw.get("alpha", 0.99), eps=kw.get("eps", 1e-8),
                                   weight_decay=kw.get("weight_decay", 0), momentum=kw.get("momentum", 0), centered=kw.get("centered", False))
    if name == "lbfgs":
        max_iter = int(kw.get("max_iter", 1))
        max_eval = kw.get("max_eval", None)
        if max_eval is None:
            max_eval = max(1, int(max_iter * 1.25))
        return torch.optim.LBFGS(params, lr=kw.get("lr", 1), max_iter=max_iter, max_eval=max_eval,
                                 tolerance_grad=kw.get("tolerance_grad", 1e-11), tolerance_change=kw.get("tolerance_change", 1e-11),
                                 history_size=kw.get("history_size", 5), line_search_fn=kw.get("line_search_fn", "strong_wolfe"))
    raise ValueError(f"Unsupported optimizer: {optimizer_type}. Choose 'adam', 'sgd', 'rmsprop', or 'lbfgs'.")


# --------------------------------------------------------------------------------------
# synthetic movies (shared by tests / bench so every arm sees the same data)
# --------------------------------------------------------------------------------------


def synthetic_movie(t: int, h: int, w: int, seed: int = 0, noise: float = 1.0, drift: float = 6.0,
                    local: float = 1.5, sigma_f: float = 0.08, integer_shifts: bool = False):
    """Band-limited random specimen + smooth global drift + smooth local field + Gaussian noise.

    Returns ``(movie (t,h,w) float32, global_shifts (t,2) px)``.  SURVEY.md §8(d).
    """
    g = torch.Generator().manual_seed(seed)
    pad = 64
    H, W = h + 2 * pad, w + 2 * pad
    white = torch.randn((H, W), generator=g)
    fy = torch.fft.fftfreq(H)[:, None]
    fx = torch.fft.rfftfreq(W)[None, :]
    lp = torch.exp(-(fy**2 + fx**2) / (2 * sigma_f**2))
    specimen = torch.fft.irfftn(torch.fft.rfftn(white) * lp, s=(H, W))
    specimen = specimen / specimen.std()
    steps = torch.randn((t, 2), generator=g)
    walk = torch.cumsum(steps, dim=0)
    walk = walk - walk[t // 2]
    walk = walk / max(float(walk.abs().max()), 1e-6) * drift
    if integer_shifts:
        walk = torch.round(walk)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    coef = torch.randn((2, 3, 4, 4), generator=g) * (0.0 if integer_shifts else local)
    frames = []
    for k in range(t):
        if integer_shifts:
            sy, sx = int(walk[k, 0]), int(walk[k, 1])
            fr = specimen[pad - sy : pad - sy + h, pad - sx : pad - sx + w]
        else:
            tyx = torch.stack([torch.full_like(yy, k / max(t - 1, 1)), yy / (h - 1), xx / (w - 1)], dim=-1)
            loc = deps.evaluate_cubic_grid_3d(coef, tyx, deps.BSPLINE_MATRIX)
            cy = yy + pad - walk[k, 0] - loc[..., 0]
            cx = xx + pad - walk[k, 1] - loc[..., 1]
            fr = deps.sample_image_2d(specimen, torch.stack([cy, cx], dim=-1), interpolation="bicubic")
        frames.append(fr + noise * torch.randn((h, w), generator=g))
    return torch.stack(frames).float().contiguous(), walk


def synthetic_movie_large(t: int, h: int, w: int, seed: int = 0, noise: float = 1.0, drift: float = 6.0,
                          local: float = 1.5, sigma_f: float = 0.08):
    """Same recipe as :func:`synthetic_movie` at a cost that suits the benchmark-sized parity cases (seconds, not
    minutes, for 40 x 2048^2): the smooth local field is evaluated on a coarse 33 x 33 lattice and brought to pixel
    resolution with ``F.interpolate`` (bicubic), and frames are resampled with ``F.grid_sample`` (bicubic) directly.

    Returns ``(movie (t,h,w) float32, global_shifts (t,2) px)``.  Used by ``tests/golden/make_golden.py`` and by the
    ``-m gpu`` tests that regenerate the same movie from its seed."""
    g = torch.Generator().manual_seed(seed)
    pad = 64
    H, W = h + 2 * pad, w + 2 * pad
    white = torch.randn((H, W), generator=g)
    fy = torch.fft.fftfreq(H)[:, None]
    fx = torch.fft.rfftfreq(W)[None, :]
    lp = torch.exp(-(fy**2 + fx**2) / (2 * sigma_f**2))
    specimen = torch.fft.irfftn(torch.fft.rfftn(white) * lp, s=(H, W))
    specimen = (specimen / specimen.std()).float()
    steps = torch.randn((t, 2), generator=g)
    walk = torch.cumsum(steps, dim=0)
    walk = walk - walk[t // 2]
    walk = walk / max(float(walk.abs().max()), 1e-6) * drift
    coef = torch.randn((2, 3, 4, 4), generator=g) * local
    n = 33
    ly, lx = torch.meshgrid(torch.linspace(0, 1, n), torch.linspace(0, 1, n), indexing="ij")
    yy = torch.arange(h, dtype=torch.float32)[:, None]
    xx = torch.arange(w, dtype=torch.float32)[None, :]
    frames = torch.empty((t, h, w), dtype=torch.float32)
    for k in range(t):
        tyx = torch.stack([torch.full_like(ly, k / max(t - 1, 1)), ly, lx], dim=-1)
        coarse = deps.evaluate_cubic_grid_3d(coef, tyx, deps.BSPLINE_MATRIX)  # (n, n, 2)
        loc = F.interpolate(coarse.permute(2, 0, 1)[None], size=(h, w), mode="bicubic", align_corners=True)[0]
        cy = yy + (pad - float(walk[k, 0])) - loc[0]
        cx = xx + (pad - float(walk[k, 1])) - loc[1]
        grid = torch.stack([cx / (W - 1) * 2 - 1, cy / (H - 1) * 2 - 1], dim=-1)[None]
        fr = F.grid_sample(specimen[None, None], grid, mode="bicubic", padding_mode="border", align_corners=True)[0, 0]
        frames[k] = fr + noise * torch.randn((h, w), generator=g)
    return frames.contiguous(), walk


# --------------------------------------------------------------------------------------
# movie preparation   (examples/ttMotion.py:90-202 -- example code, not part of the package)
# --------------------------------------------------------------------------------------


def _hash3(a, b, c):
    """Position hash of csrc/prepare.cu (uint32 arithmetic)."""
    a, b, c = (np.asarray(v, dtype=np.uint64) for v in (a, b, c))
    m = np.uint64(0xFFFFFFFF)
    h = ((a * np.uint64(0x9E3779B1)) & m) ^ (((b + np.uint64(0x7F4A7C15)) & m) * np.uint64(0x85EBCA77) & m) ^ (
        ((c + np.uint64(0x165667B1)) & m) * np.uint64(0xC2B2AE3D) & m)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x2C1B3C6D)) & m
    h ^= h >> np.uint64(12)
    h = (h * np.uint64(0x297A2D39)) & m
    h ^= h >> np.uint64(15)
    return h


def prepare_movie(movie: np.ndarray, gain: np.ndarray | None = None, hot_pixel_threshold: float | None = None,
                  zero_frame_means: bool = False):
    """``gain_correct`` (:85-123, the multiply), ``remove_hot_pixels`` (:125-178) and ``set_frames_mean_zero`` (:180-202)
    in the order of the example's ``main`` (:372-381), on a float32 copy of the movie.

    One documented deviation: the reference replaces a hot pixel by ``np.random.choice`` of its neighbours, sequentially
    and in place; here (and in csrc/prepare.cu) the neighbour is picked by a hash of (frame, y, x) and read from the frame
    as it was before any replacement -- the same distribution, reproducible, order independent.  Returns
    ``(movie float32, number of hot pixels)``."""
    out = movie.astype(np.float32)
    if gain is not None:
        out = out * gain.astype(np.float32)
    n_hot = 0
    if hot_pixel_threshold is not None:
        t, h, w = out.shape
        fixed = out.copy()
        for f in range(t):
            frame = out[f]
            mean, std = np.mean(frame, dtype=np.float64), np.std(frame, dtype=np.float64)
            lim = np.float32(hot_pixel_threshold) * np.float32(std)
            hot = (frame > np.float32(mean) + lim) | (frame < np.float32(mean) - lim)
            for y, x in zip(*np.where(hot)):
                y0, y1, x0, x1 = max(0, y - 1), min(h - 1, y + 1), max(0, x - 1), min(w - 1, x + 1)
                cols, cells = x1 - x0 + 1, (y1 - y0 + 1) * (x1 - x0 + 1)
                if cells <= 1:
                    continue
                pick = int(_hash3(f, y, x) % np.uint64(cells - 1))
                if pick >= (y - y0) * cols + (x - x0):
                    pick += 1
                fixed[f, y, x] = frame[y0 + pick // cols, x0 + pick % cols]
                n_hot += 1
        out = fixed
    if zero_frame_means:
        out = out - np.mean(out, axis=(1, 2), keepdims=True, dtype=np.float64).astype(np.float32)
    return out, n_hot


# --------------------------------------------------------------------------------------
# iterative refinement driver   (examples/ttMotion.py:287-329 -- example code, sketched there)
# --------------------------------------------------------------------------------------


def estimate_motion_pipeline(image, pixel_spacing: float, patch_sidelength: int, frequency_range=(300, 10), b_factor: float = 500,
                             n_refinements: int = 0, refinement_tolerance: float = 1e-3, smoothing_window_size: int = 5):
    """The B200 build's ``estimate_motion`` restated with the functions above: global estimate (whole pixels) -> patch
    cross-correlation on the rigidly pre-shifted movie (global / pixel_spacing handed over as the reference's
    correct_motion_fast route wants pixels, quirk Q2) WITHOUT smoothing -> the base it accumulated on swapped for the
    true global field -> Savitzky-Golay + one joint mean; then up to ``n_refinements`` passes of the example's loop
    (``deformation_field=`` the cumulative field of the pass before) until the mean absolute change drops below the
    tolerance.  Returns ``(field, [mean absolute change per pass])``."""
    g = estimate_global_motion(image, pixel_spacing, b_factor=b_factor, frequency_range=frequency_range)
    pre = g / pixel_spacing
    handed = pre.clone()
    f, _ = estimate_motion_cross_correlation_patches(image, pixel_spacing, patch_sidelength=patch_sidelength, b_factor=b_factor,
                                                     frequency_range=frequency_range, deformation_field=pre, smooth=False)
    shape = f.shape[1:]
    field = f - resample_deformation_field(-handed, shape) + resample_deformation_field(g, shape)
    field = temporal_smoothing(field, smoothing_window_size)
    field = field - field.mean()
    history = []
    for _ in range(n_refinements):
        refined, _ = estimate_motion_cross_correlation_patches(image, pixel_spacing, patch_sidelength=patch_sidelength,
                                                               b_factor=b_factor, frequency_range=frequency_range,
                                                               deformation_field=field.clone())
        history.append(float((refined - field).abs().mean()))
        field = refined
        if history[-1] < refinement_tolerance:
            break
    return field, history
