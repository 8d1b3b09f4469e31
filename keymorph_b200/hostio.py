"""Host <-> device plumbing for streams of volume pairs: the copy of item i+1 runs on a side stream
while item i is being registered (two CUDA streams, event-ordered; torch owns the memory)."""
from __future__ import annotations

import torch


def prefetch_to_device(items, device, depth=2):
    """Yield tuples of device tensors for an iterable of tuples of (pinned) host tensors.  The H2D
    copies of the NEXT item are enqueued on a dedicated copy stream before the current item is
    handed out, so they overlap with the kernels the caller launches on the current stream.

    The device side is a ring of `depth` persistent staging buffers per tuple position (allocated
    once per shape): nothing goes through the caching allocator per item, which would otherwise
    fall back to cudaMalloc -- a device-wide synchronisation -- whenever a freed block is still
    guarded by a cross-stream event (measured: 30 ms instead of 9 ms per step with two ranks per
    node).  A yielded tensor is valid until the generator has been advanced `depth - 1` more times."""
    device = torch.device(device)
    copy_stream = torch.cuda.Stream(device)
    compute = torch.cuda.current_stream(device)
    ring = [None] * depth          # per slot: tuple of device buffers
    done = [None] * depth          # per slot: event recorded when the consumer moved past that item

    def stage(item, slot):
        bufs = ring[slot]
        if bufs is None or len(bufs) != len(item) or any(
                b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, item)):
            bufs = tuple(torch.empty(t.shape, dtype=t.dtype, device=device) for t in item)
            ring[slot] = bufs
        with torch.cuda.stream(copy_stream):
            if done[slot] is not None:
                copy_stream.wait_event(done[slot])     # the kernels that read this slot have been enqueued
            for b, t in zip(bufs, item):
                b.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs, ev

    it = iter(items)
    k = 0
    try:
        nxt = stage(next(it), 0)
    except StopIteration:
        return
    while nxt is not None:
        cur, ev = nxt
        slot = k % depth
        try:
            nxt = stage(next(it), (k + 1) % depth)
        except StopIteration:
            nxt = None
        compute.wait_event(ev)
        yield cur
        # the consumer has enqueued its work on item k: the slot may be overwritten after that work
        d = torch.cuda.Event()
        d.record(compute)
        done[slot] = d
        k += 1


# --------------------------------------------------------------------------------------------
# NIfTI-1 input without nibabel / torchio (scripts/register.py:40-118 goes through torchio's
# ScalarImage / LabelMap, hyperparameters.py:4-11 TRANSFORM = ToCanonical, Mask, Resize(128),
# rescale_intensity).  Host-side Python/numpy only: file parsing is not on the GPU hot path.
_NIFTI_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4"}


def read_nifti(path):
    """Read a NIfTI-1 single-file image (.nii or .nii.gz).  Returns (array (X,Y,Z[,T]) in file
    orientation with scl_slope / scl_inter applied when set, affine (4,4) float64 from the sform,
    or the pixdim-scaled identity when sform_code == 0)."""
    import gzip
    import struct

    import numpy as np
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    for end in ("<", ">"):
        if struct.unpack(end + "i", raw[:4])[0] == 348:
            break
    else:
        raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    if raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError(f"{path}: bad NIfTI magic {raw[344:348]!r}")
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype = struct.unpack(end + "h", raw[70:72])[0]
    if datatype not in _NIFTI_DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype {datatype}")
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset = int(struct.unpack(end + "f", raw[108:112])[0])
    slope, inter = struct.unpack(end + "2f", raw[112:120])
    qform_code, sform_code = struct.unpack(end + "2h", raw[252:256])
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    dt = np.dtype(end + _NIFTI_DTYPES[datatype])
    count = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=count, offset=vox_offset).reshape(shape, order="F")
    # nibabel semantics: a zero or non-finite slope (or a non-finite intercept) means "no scaling"
    if np.isfinite(slope) and np.isfinite(inter) and slope != 0.0 and (slope != 1.0 or inter != 0.0):
        data = data.astype(np.float64) * slope + inter
    affine = np.eye(4)
    if sform_code > 0:
        affine[:3, :] = np.array(struct.unpack(end + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    elif qform_code > 0:
        # NIfTI-1 method 2: unit quaternion (b, c, d) + offsets, pixdim[0] = qfac (handedness of the k axis)
        b, c, d, qx, qy, qz = (float(v) for v in struct.unpack(end + "6f", raw[256:280]))
        a2 = 1.0 - (b * b + c * c + d * d)
        if a2 < 1e-7:                       # 180-degree rotation: renormalise (b, c, d), a = 0
            nrm = 1.0 / np.sqrt(b * b + c * c + d * d)
            b, c, d, a = b * nrm, c * nrm, d * nrm, 0.0
        else:
            a = np.sqrt(a2)
        R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                      [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                      [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        affine[:3, :3] = R * np.array([pixdim[1], pixdim[2], pixdim[3] * qfac], dtype=np.float64)[None, :]
        affine[:3, 3] = (qx, qy, qz)
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    return data, affine


def to_canonical(data, affine):
    """torchio.ToCanonical / nibabel.as_closest_canonical for (near) axis-aligned affines: permute
    and flip the voxel axes so that they run left->Right, posterior->Anterior, inferior->Superior."""
    import numpy as np
    R = affine[:3, :3]
    perm = [int(np.argmax(np.abs(R[i, :]))) for i in range(3)]   # voxel axis that drives world axis i
    if sorted(perm) != [0, 1, 2]:
        raise ValueError("to_canonical: oblique affine; resample first")
    data = np.transpose(data, perm + list(range(3, data.ndim)))
    aff = affine[:, perm + [3]].copy()
    for i in range(3):
        if aff[i, i] < 0:
            data = np.flip(data, axis=i)
            aff[:3, 3] += aff[:3, i] * (data.shape[i] - 1)
            aff[:3, i] = -aff[:3, i]
    return np.ascontiguousarray(data), aff


def load_volume(path, size=None, labels=False):
    """One subject as the reference's loaders deliver it to the model: canonical orientation, resized
    to `size`^3 when the volume is an integer multiple of it (block mean for images, strided pick
    for label maps -- a deterministic stand-in for torchio.Resize), intensities
    rescaled to [0, 1] (keymorph/utils.py:78-94).  Returns (tensor (1,1,D,H,W) float32 | uint8, affine)."""
    import numpy as np
    data, affine = to_canonical(*read_nifti(path))
    if data.ndim == 4:
        data = data[..., 0]
    if size is not None and data.shape != (size,) * 3:
        f = [s // size for s in data.shape]
        if any(s != size * k or k < 1 for s, k in zip(data.shape, f)):
            raise ValueError(f"load_volume: {data.shape} is not an integer multiple of {size}")
        old = affine
        affine = affine.copy()
        if labels:
            data = data[::f[0], ::f[1], ::f[2]]         # coarse voxel 0 IS fine voxel 0: origin unchanged
        else:
            data = data.astype(np.float32).reshape(size, f[0], size, f[1], size, f[2]).mean(axis=(1, 3, 5))
            # the centre of coarse voxel 0 sits at fine index (f - 1) / 2 along every axis
            affine[:3, 3] = old[:3, 3] + old[:3, :3] @ ((np.array(f, dtype=np.float64) - 1.0) / 2.0)
        affine[:3, :3] = old[:3, :3] * np.array(f, dtype=np.float64)[None, :]
    if labels:
        if data.size and (data.max() > 255 or data.min() < 0):
            raise ValueError(f"load_volume: label ids span [{data.min()}, {data.max()}]; the label-map kernels "
                             "take uint8 ids (< 256): remap the labels to contiguous ids first")
        t = torch.from_numpy(np.ascontiguousarray(data).astype(np.uint8))
    else:
        a = np.ascontiguousarray(data).astype(np.float32)
        lo, hi = float(a.min()), float(a.max())
        a = (a - lo) / (hi - lo) if hi > lo else np.zeros_like(a)
        t = torch.from_numpy(a)
    return t[None, None], affine
