"""Seeded synthetic inputs of the shapes BASELINE.json names: there is no network for the
reference's lib.mil / big-lib.mil, so benchmarks and parity tests use these. numpy only."""
import numpy as np


def make_library(n: int, size: int, seed: int = 1004, out: np.ndarray | None = None) -> np.ndarray:
    """n x size x size x 3 uint8 BGR: smooth random fields (8x8 noise, nearest-upsampled, plus fine noise) around
    per-image mean colours spread over the RGB cube; every 97th image is a near-duplicate of its predecessor so that
    a tail of near-ties exists for the tie band."""
    rng = np.random.default_rng(seed)
    lib = out if out is not None else np.empty((n, size, size, 3), np.uint8)
    low_n = 8 if size >= 8 else 1
    rep = -(-size // low_n)
    chunk = 256
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        mean = rng.integers(16, 240, (m, 1, 1, 3)).astype(np.float32)
        low = rng.normal(0, 28, (m, low_n, low_n, 3)).astype(np.float32)
        up = low.repeat(rep, 1).repeat(rep, 2)[:, :size, :size]
        fine = rng.normal(0, 6, (m, size, size, 3)).astype(np.float32)
        lib[s:s + m] = np.clip(mean + up + fine, 0, 255).astype(np.uint8)
    for i in range(97, n, 97):
        d = lib[i - 1].astype(np.int16)
        d[::7, ::5] += 1
        lib[i] = np.clip(d, 0, 255).astype(np.uint8)
    return lib


def make_main_image(h: int, w: int, seed: int = 2004, noise_fraction: float = 0.4, block: int = 64, knot: int = 0) -> np.ndarray:
    """h x w x 3 uint8 BGR: smooth colour field with uniform-noise blocks over ~noise_fraction of the area, so that the
    reference's 5.6-bit entropy rule splits a known share of cells."""
    rng = np.random.default_rng(seed)
    knot = knot or 8 * block  # colour knots are far apart so that smooth cells stay below the entropy threshold
    kh, kw = -(-h // knot) + 1, -(-w // knot) + 1
    gh, gw = -(-h // block) + 1, -(-w // block) + 1
    knots = rng.integers(0, 256, (kh, kw, 3)).astype(np.float32)
    ys = np.arange(h, dtype=np.float32) / knot
    xs = np.arange(w, dtype=np.float32) / knot
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    img = np.empty((h, w, 3), np.float32)
    rows = 512
    for r in range(0, h, rows):  # bilinear interpolation of the knot colours, in row slabs
        yy = y0[r:r + rows]
        top = knots[yy][:, x0] * (1 - fx) + knots[yy][:, x0 + 1] * fx
        bot = knots[yy + 1][:, x0] * (1 - fx) + knots[yy + 1][:, x0 + 1] * fx
        img[r:r + rows] = top * (1 - fy[r:r + rows]) + bot * fy[r:r + rows]
    noisy = rng.random((gh, gw)) < noise_fraction
    mask = noisy.repeat(block, 0).repeat(block, 1)[:h, :w]
    noise = rng.integers(0, 256, (h, w, 3)).astype(np.float32)
    img = np.where(mask[..., None], noise, img)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def hexagon_mask(size: int) -> np.ndarray:
    """A pointy-top hexagon mask (0/255), stand-in for Cells/Hexagon.mcs where the reference checkout is absent."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float64)
    cx = cy = (size - 1) / 2
    dx, dy = np.abs(x - cx) / (size / 2), np.abs(y - cy) / (size / 2)
    inside = (dx <= 0.87) & (dy <= 1.0) & (dy + dx * 0.5774 <= 1.0)
    return np.where(inside, 255, 0).astype(np.uint8)


def triangle_mask(size: int) -> np.ndarray:
    """An isosceles triangle (apex up): with alternate-column vertical flips it tiles like IsocelesTriangle.mcs."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float64)
    half = np.abs(x - (size - 1) / 2)
    return np.where(half <= (y + 1) / 2, 255, 0).astype(np.uint8)


def make_photo_library(src: np.ndarray, n: int, size: int, seed: int = 2002) -> np.ndarray:
    """n x size x size x 3 uint8 BGR library cut from a photograph: seeded square crops of side k * size (k = 1..6) at random
    positions, reduced by the k x k block mean. Substitute for the reference's Library/lib.mil / big-lib.mil, which are missing
    from its checkout ("build a deterministic one from tiles of the sample images")."""
    rng = np.random.default_rng(seed)
    h, w = src.shape[:2]
    kmax = max(1, min(6, min(h, w) // size))
    lib = np.empty((n, size, size, 3), np.uint8)
    for i in range(n):
        k = int(rng.integers(1, kmax + 1))
        side = k * size
        y, x = int(rng.integers(0, h - side + 1)), int(rng.integers(0, w - side + 1))
        crop = src[y:y + side, x:x + side].astype(np.uint32).reshape(size, k, size, k, 3)
        lib[i] = ((crop.sum((1, 3)) + (k * k) // 2) // (k * k)).astype(np.uint8)
    return lib
