"""MRC / TIFF I/O (replacement of mrcfile / skimage.io+tifffile, src/flowdenoising.py:466-475, :539-548) and the CLI
surface (:384-415). CPU only."""
import struct

import numpy as np
import pytest

from flowdenoising_b200 import volume_io


@pytest.mark.parametrize("dtype,mode", [(np.float32, 2), (np.int8, 0), (np.int16, 1), (np.uint16, 6), (np.float16, 12)])
def test_mrc_read_modes(tmp_path, dtype, mode):
    rng = np.random.default_rng(0)
    vol = (rng.random((3, 5, 7)) * 100).astype(dtype)
    hdr = bytearray(1024)
    struct.pack_into("<4i", hdr, 0, 7, 5, 3, mode)
    struct.pack_into("<i", hdr, 92, 64)            # extended header of 64 bytes must be skipped
    hdr[208:212] = b"MAP "
    hdr[212:216] = b"\x44\x44\x00\x00"
    p = tmp_path / "v.mrc"
    p.write_bytes(bytes(hdr) + b"\0" * 64 + vol.tobytes())
    got = volume_io.read_volume(str(p))
    assert got.dtype == np.dtype(dtype) and got.shape == (3, 5, 7) and np.array_equal(got, vol)
    mm = volume_io.read_mrc(str(p), memory_map=True)
    assert np.array_equal(np.asarray(mm), vol)


def test_mrc_big_endian_and_roundtrip(tmp_path):
    vol = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    hdr = bytearray(1024)
    struct.pack_into(">4i", hdr, 0, 4, 3, 2, 2)
    hdr[208:212] = b"MAP "
    hdr[212:216] = b"\x11\x11\x00\x00"
    p = tmp_path / "be.MRC"
    p.write_bytes(bytes(hdr) + vol.astype(">f4").tobytes())
    assert np.array_equal(volume_io.read_volume(str(p)), vol)
    out = tmp_path / "out.mrc"
    volume_io.write_volume(str(out), vol * 2)
    back = volume_io.read_volume(str(out))
    assert back.dtype == np.float32 and np.array_equal(back, vol * 2)
    raw = out.read_bytes()
    assert struct.unpack_from("<4i", raw, 0) == (4, 3, 2, 2) and raw[208:212] == b"MAP "
    assert abs(struct.unpack_from("<f", raw, 84)[0] - float((vol * 2).mean())) < 1e-5    # dmean


def test_mrc_errors(tmp_path):
    p = tmp_path / "bad.mrc"
    p.write_bytes(b"\0" * 100)
    with pytest.raises(ValueError):
        volume_io.read_volume(str(p))
    with pytest.raises(FileNotFoundError):
        volume_io.read_volume(str(tmp_path / "missing.mrc"))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_tiff_roundtrip_and_pillow_compat(tmp_path, dtype):
    rng = np.random.default_rng(1)
    vol = (rng.random((4, 6, 9)) * 200).astype(dtype)
    p = tmp_path / "v.tif"
    volume_io.write_volume(str(p), vol)                 # always float32, like the reference (:548)
    back = volume_io.read_volume(str(p))
    assert back.dtype == np.float32 and np.array_equal(back, vol.astype(np.float32))
    Image = pytest.importorskip("PIL.Image")
    with Image.open(str(p)) as im:                       # an independent reader agrees
        assert im.n_frames == 4
        im.seek(2)
        assert np.array_equal(np.array(im), vol[2].astype(np.float32))
    # a TIFF written by an independent writer is read back (uint8/uint16 pages -> float32 like imread(...).astype)
    if dtype != np.float32:
        q = tmp_path / "pil.tif"
        frames = [Image.fromarray(vol[i]) for i in range(vol.shape[0])]
        frames[0].save(str(q), save_all=True, append_images=frames[1:])
        got = volume_io.read_volume(str(q))
        assert got.dtype == np.float32 and np.array_equal(got, vol.astype(np.float32))


def test_extension_rules_match_reference():
    assert volume_io.is_mrc_input("a.mrc") and volume_io.is_mrc_input("a.MRCS") and not volume_io.is_mrc_input("a.tif")
    assert volume_io.is_mrc_output("a.mrc") and volume_io.is_mrc_output("a.MRC") and not volume_io.is_mrc_output("a.mrcs")


def test_cli_parser_matches_reference_flags():
    from flowdenoising_b200 import flowdenoising as fd
    a = fd.parser.parse_args([])
    assert a.input == "./volume.mrc" and a.output == "./denoised_volume.mrc"
    assert tuple(a.sigma) == (2.0, 2.0, 2.0) and a.levels == 3 and a.winsize == 5 and a.verbosity == 0
    assert not a.no_OF and not a.memory_map and not a.recompute_flow and not a.show_fingerprint
    a = fd.parser.parse_args("-i x.tif -o y.mrc -s 4 2 2 -l 5 -w 9 -v 2 -n -m -p 3 --recompute_flow".split())
    assert (a.input, a.output, a.levels, a.winsize, a.verbosity, a.number_of_processes) == ("x.tif", "y.mrc", 5, 9, 2, 3)
    assert [float(s) for s in a.sigma] == [4.0, 2.0, 2.0] and a.no_OF and a.memory_map and a.recompute_flow
    assert (fd.OF_LEVELS, fd.OF_WINDOW_SIZE, fd.OF_ITERS, fd.OF_POLY_N, fd.OF_POLY_SIGMA, fd.SIGMA) == (3, 5, 3, 5, 1.2, 2.0)
    assert a.iterations == 3 and a.poly_n == 5 and a.poly_sigma == 1.2


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flowdenoising_b200 import flowdenoising as fd
    from flowdenoising_b200._lib import FdnError
    vol = np.zeros((4, 8, 8), np.float32)
    with pytest.raises(FdnError):
        fd.GaussianDenoising(1, vol).filter([np.ones(1)] * 3)
    with pytest.raises(FdnError):
        fd.warp_slice(vol[0], np.zeros((8, 8, 2), np.float32))
