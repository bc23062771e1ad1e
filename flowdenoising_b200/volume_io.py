"""MRC2014 and multi-page TIFF volume I/O in pure NumPy.

Replaces the reference's use of ``mrcfile`` / ``skimage.io`` + ``tifffile``
(/root/reference/src/flowdenoising.py:466-475 read, :539-548 write), none of which is installed here.
Volumes are [Z, Y, X] arrays (x fastest), exactly what ``mrcfile.open(...).data`` / ``skimage.io.imread`` return.

MRC: modes 0 (int8), 1 (int16), 2 (float32), 6 (uint16), 12 (float16) in; mode 2 out (:544).
TIFF: uncompressed grayscale u8/i8/u16/i16/u32/i32/f32/f64 strips, classic or BigTIFF, either byte order; other
TIFFs (compressed, tiled) go through Pillow if it is importable. Output is float32 (:548), BigTIFF above 4 GiB.
"""
from __future__ import annotations

import os
import struct
import time

import numpy as np

_MRC_MODES = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16, 12: np.float16}


def is_mrc_input(path: str) -> bool:
    """Same test as the reference (:466): 'mrc' anywhere in the extension, case-insensitive."""
    return "mrc" in path.split('.')[-1].lower()


def is_mrc_output(path: str) -> bool:
    """Same test as the reference (:539): extension exactly 'mrc' or 'MRC'."""
    ext = path.split('.')[-1]
    return ext == "MRC" or ext == "mrc"


# ---------------------------------------------------------------- MRC
def read_mrc(path: str, memory_map: bool = False) -> np.ndarray:
    with open(path, "rb") as f:
        header = f.read(1024)
        if len(header) < 1024:
            raise ValueError(f"{path}: not an MRC file (short header)")
        machst = header[212:214]
        if machst[:1] in (b"\x44", b"\x41"):
            bo = "<"
        elif machst[:1] == b"\x11":
            bo = ">"
        else:  # unknown stamp: pick the byte order that makes nx/ny/nz/mode sane
            bo = "<"
            nx, ny, nz, mode = struct.unpack("<4i", header[:16])
            if not (0 < nx < 1 << 20 and 0 < ny < 1 << 20 and 0 < nz < 1 << 20 and 0 <= mode < 200):
                bo = ">"
        nx, ny, nz, mode = struct.unpack(bo + "4i", header[:16])
        nsymbt = struct.unpack(bo + "i", header[92:96])[0]
        if mode not in _MRC_MODES:
            raise ValueError(f"{path}: unsupported MRC mode {mode}")
        if nx <= 0 or ny <= 0 or nz <= 0 or nsymbt < 0:
            raise ValueError(f"{path}: corrupt MRC header")
        dt = np.dtype(_MRC_MODES[mode]).newbyteorder(bo)
        offset = 1024 + nsymbt
        count = nx * ny * nz
        if memory_map:
            return np.memmap(path, dtype=dt, mode="r", offset=offset, shape=(nz, ny, nx))
        f.seek(offset)
        data = np.fromfile(f, dtype=dt, count=count)
        if data.size != count:
            raise ValueError(f"{path}: truncated MRC data block")
    data = data.reshape(nz, ny, nx)
    if not dt.isnative:
        data = data.astype(dt.newbyteorder("="))
    return data


def _mrc_header(nz, ny, nx, dmin, dmax, dmean, rms):
    h = bytearray(1024)
    struct.pack_into("<4i", h, 0, nx, ny, nz, 2)
    struct.pack_into("<3i", h, 16, 0, 0, 0)
    struct.pack_into("<3i", h, 28, nx, ny, nz)
    struct.pack_into("<3f", h, 40, float(nx), float(ny), float(nz))   # 1 A voxels
    struct.pack_into("<3f", h, 52, 90.0, 90.0, 90.0)
    struct.pack_into("<3i", h, 64, 1, 2, 3)
    struct.pack_into("<3f", h, 76, dmin, dmax, dmean)
    struct.pack_into("<2i", h, 88, 1, 0)           # ispg = 1 (volume), nsymbt = 0
    h[104:108] = b"\0\0\0\0"                       # exttyp
    struct.pack_into("<i", h, 108, 20140)          # nversion
    h[208:212] = b"MAP "
    h[212:216] = b"\x44\x44\x00\x00"
    struct.pack_into("<f", h, 216, rms)
    label = ("Created by flowdenoising_b200 " + time.strftime("%Y-%m-%d %H:%M:%S")).encode()[:80]
    struct.pack_into("<i", h, 220, 1)
    h[224:224 + len(label)] = label
    return h


def _stats(vol):
    """min, max, mean, rms deviation, plane by plane (works on memory maps without materialising the volume)."""
    dmin, dmax, s1, s2 = np.inf, -np.inf, 0.0, 0.0
    for z in range(vol.shape[0]):
        p = np.asarray(vol[z], dtype=np.float64)
        dmin = min(dmin, float(p.min())); dmax = max(dmax, float(p.max()))
        s1 += float(p.sum()); s2 += float(np.square(p).sum())
    n = float(vol.size)
    mean = s1 / n
    return dmin, dmax, mean, float(np.sqrt(max(0.0, s2 / n - mean * mean)))


def write_mrc(path: str, vol: np.ndarray):
    vol = np.ascontiguousarray(vol, dtype="<f4")
    if vol.ndim != 3:
        raise ValueError("MRC output needs a 3-D volume")
    nz, ny, nx = vol.shape
    h = _mrc_header(nz, ny, nx, *_stats(vol))
    with open(path, "wb") as f:
        f.write(h)
        vol.tofile(f)


def create_mrc_memmap(path: str, shape) -> np.memmap:
    """A mode-2 (float32) MRC file of the given (nz, ny, nx) shape whose data block is returned as a writable memory
    map: out-of-core results are written in place (`-m`). Call finish_mrc_memmap() when the data is complete."""
    nz, ny, nx = (int(v) for v in shape)
    with open(path, "wb") as f:
        f.write(_mrc_header(nz, ny, nx, 0.0, 0.0, 0.0, 0.0))
        f.truncate(1024 + 4 * nz * ny * nx)
    return np.memmap(path, dtype="<f4", mode="r+", offset=1024, shape=(nz, ny, nx))


def finish_mrc_memmap(path: str, mm: np.memmap):
    """Flushes the map and fills in the header's density statistics."""
    mm.flush()
    nz, ny, nx = mm.shape
    h = _mrc_header(nz, ny, nx, *_stats(mm))
    with open(path, "r+b") as f:
        f.write(h)


# ---------------------------------------------------------------- TIFF
_TIFF_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
               16: "Q", 17: "q", 18: "Q"}


def _tiff_pages(buf, path):
    bo = {b"II": "<", b"MM": ">"}.get(bytes(buf[:2]))
    if bo is None:
        raise ValueError(f"{path}: not a TIFF file")
    magic = struct.unpack_from(bo + "H", buf, 2)[0]
    if magic == 42:
        big = False
        ifd = struct.unpack_from(bo + "I", buf, 4)[0]
    elif magic == 43:
        big = True
        ifd = struct.unpack_from(bo + "Q", buf, 8)[0]
    else:
        raise ValueError(f"{path}: bad TIFF magic {magic}")
    pages = []
    seen = set()
    while ifd and ifd not in seen:
        seen.add(ifd)
        if big:
            n = struct.unpack_from(bo + "Q", buf, ifd)[0]; p = ifd + 8; esz = 20; inl = 8
        else:
            n = struct.unpack_from(bo + "H", buf, ifd)[0]; p = ifd + 2; esz = 12; inl = 4
        tags = {}
        for i in range(n):
            e = p + i * esz
            tag, typ = struct.unpack_from(bo + "HH", buf, e)
            cnt = struct.unpack_from(bo + ("Q" if big else "I"), buf, e + 4)[0]
            fmt = _TIFF_TYPES.get(typ)
            if fmt is None:
                continue
            per = struct.calcsize("=" + fmt)
            voff = e + (12 if big else 8)
            if per * cnt > inl:
                voff = struct.unpack_from(bo + ("Q" if big else "I"), buf, voff)[0]
            if typ == 2:
                tags[tag] = bytes(buf[voff:voff + cnt])
            else:
                nvals = cnt * len(fmt)
                tags[tag] = struct.unpack_from(bo + str(nvals) + fmt[0], buf, voff)
        pages.append(tags)
        ifd = struct.unpack_from(bo + ("Q" if big else "I"), buf, p + n * esz)[0]
    return bo, pages


def read_tiff(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 8:
        raise ValueError(f"{path}: not a TIFF file")
    try:
        bo, pages = _tiff_pages(buf, path)
        out = []
        for t in pages:
            comp = t.get(259, (1,))[0]
            spp = t.get(277, (1,))[0]
            if comp != 1 or spp != 1 or 322 in t:
                raise NotImplementedError
            w, h = t[256][0], t[257][0]
            bits = t.get(258, (1,))[0]
            fmt = t.get(339, (1,))[0]
            kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
            if kind is None or bits not in (8, 16, 32, 64):
                raise NotImplementedError
            dt = np.dtype(f"{bo}{kind}{bits // 8}")
            offs, cnts = t[273], t[279]
            if len(offs) == 1:
                page = np.frombuffer(buf, dtype=dt, count=w * h, offset=offs[0])
            else:
                page = np.concatenate([np.frombuffer(buf, dtype=np.uint8, count=c, offset=o)
                                       for o, c in zip(offs, cnts)]).view(dt)[:w * h]
            out.append(page.reshape(h, w))
        if not out:
            raise ValueError(f"{path}: no images")
        vol = np.stack(out) if len(out) > 1 else out[0][None]
        return vol.astype(vol.dtype.newbyteorder("="))
    except NotImplementedError:
        pass
    # compressed / tiled / palette TIFFs: Pillow
    from PIL import Image
    with Image.open(path) as im:
        frames = []
        for i in range(getattr(im, "n_frames", 1)):
            im.seek(i)
            frames.append(np.array(im))
    return np.stack(frames)


def write_tiff(path: str, vol: np.ndarray):
    """Multi-page uncompressed float32 grayscale; one strip per page; BigTIFF when the file exceeds 4 GiB."""
    vol = np.ascontiguousarray(vol, dtype="<f4")
    if vol.ndim == 2:
        vol = vol[None]
    nz, ny, nx = vol.shape
    page_bytes = ny * nx * 4
    big = nz * (page_bytes + 256) + 16 > (1 << 32) - 1
    desc = ('{"shape": [%d, %d, %d]}' % (nz, ny, nx)).encode() + b"\0"
    with open(path, "wb") as f:
        if big:
            f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, 16))
        else:
            f.write(struct.pack("<2sHI", b"II", 42, 8))
        pos = f.tell()
        for z in range(nz):
            entries = [(256, 4, nx), (257, 4, ny), (258, 3, 32), (259, 3, 1), (262, 3, 1), (273, None, None),
                       (277, 3, 1), (278, 4, ny), (279, None, page_bytes), (339, 3, 3)]
            if z == 0:
                entries.insert(5, (270, 2, desc))
            n = len(entries)
            if big:
                ifd_size = 8 + n * 20 + 8
            else:
                ifd_size = 2 + n * 12 + 4
            extra_off = pos + ifd_size
            extra = desc if z == 0 else b""
            if len(extra) % 2:
                extra += b"\0"
            data_off = extra_off + len(extra)
            next_ifd = data_off + page_bytes if z + 1 < nz else 0
            out = bytearray()
            out += struct.pack("<Q" if big else "<H", n)
            for tag, typ, val in entries:
                if tag == 273:
                    typ, val = (16 if big else 4), data_off
                if tag == 279:
                    typ = 16 if big else 4
                if typ == 2:
                    cnt = len(desc)
                    out += struct.pack("<HH", tag, 2) + struct.pack("<Q" if big else "<I", cnt)
                    out += struct.pack("<Q" if big else "<I", extra_off)
                    continue
                out += struct.pack("<HH", tag, typ) + struct.pack("<Q" if big else "<I", 1)
                code = {3: "H", 4: "I", 16: "Q"}[typ]
                v = struct.pack("<" + code, val)
                out += v + b"\0" * ((8 if big else 4) - len(v))
            out += struct.pack("<Q" if big else "<I", next_ifd)
            f.write(out)
            f.write(extra)
            vol[z].tofile(f)
            pos = data_off + page_bytes


# ---------------------------------------------------------------- front door
def read_volume(path: str, memory_map: bool = False) -> np.ndarray:
    """:466-475: MRC -> array in the file's dtype; TIFF -> float32."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if is_mrc_input(path):
        return read_mrc(path, memory_map=memory_map)
    return read_tiff(path).astype(np.float32)


def write_volume(path: str, vol: np.ndarray):
    """:539-548: float32 MRC if the extension is exactly mrc/MRC, else float32 TIFF."""
    if is_mrc_output(path):
        write_mrc(path, vol)
    else:
        write_tiff(path, vol)
