"""End-to-end CLI drop-in on the GPU: MRC/TIFF in -> flowdenoising_b200.flowdenoising.main -> MRC/TIFF out, compared
with the reference's outputs (tests/golden). Mirrors how the reference is run (src/test_me.sh:5-6)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def test_cli_mrc_and_tiff_roundtrip(tmp_path, golden):
    from flowdenoising_b200 import flowdenoising as fd
    from flowdenoising_b200 import volume_io
    g = golden("toy_of.npz")
    vol = g["vol"].astype(np.float32)
    sig = [str(float(s)) for s in g["sigmas"]]
    src = tmp_path / "toy.mrc"
    volume_io.write_mrc(str(src), vol)
    out = tmp_path / "out.mrc"
    assert fd.main(["-i", str(src), "-o", str(out), "-s", *sig, "-l", "3", "-w", "5"]) == 0
    assert np.array_equal(volume_io.read_volume(str(out)), g["ZYX"])          # full Z+Y+X result
    # the reference CLI writes the Z+Y intermediate (quirk Q1, src/flowdenoising.py:520): --compat_zy_output
    out2 = tmp_path / "out_zy.tif"
    assert fd.main(["-i", str(src), "-o", str(out2), "-s", *sig, "--compat_zy_output"]) == 0
    assert np.array_equal(volume_io.read_volume(str(out2)), g["ZY"])
    # uint8 TIFF input, OF disabled (cfg 4 style input type, src/flowdenoising.py:475)
    n = golden("toy_noof.npz")
    tif = tmp_path / "toy_u8.tif"
    from PIL import Image
    frames = [Image.fromarray(n["vol"][i]) for i in range(n["vol"].shape[0])]
    frames[0].save(str(tif), save_all=True, append_images=frames[1:])
    out3 = tmp_path / "gauss.mrc"
    assert fd.main(["-i", str(tif), "-o", str(out3), "-s", *sig, "-n"]) == 0
    assert np.array_equal(volume_io.read_volume(str(out3)), n["ZYX"])
    # --recompute_flow
    r = golden("toy_of_recompute.npz")
    out4 = tmp_path / "rec.mrc"
    assert fd.main(["-i", str(src), "-o", str(out4), "-s", sig[0], "0.1", "0.1", "--recompute_flow",
                    "--compat_zy_output"]) == 0   # sigma 0.1 -> r = int(0.9) = 0: single-tap kernels, the Y pass is the identity
    assert np.array_equal(volume_io.read_volume(str(out4)), r["Z"])
