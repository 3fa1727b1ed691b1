"""fpie_b200.io against the contract of the reference's fpie/io.py:10-39: same arrays
for colour / grey / RGBA files, the same error for an unreadable file, the same default mask, and an
asynchronous writer whose files equal cv2.imwrite's.  No GPU needed: without a CUDA runtime the staging
buffers are ordinary arrays."""

import os
import warnings

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from fpie_b200 import io as bio  # noqa: E402

REF = os.environ.get("FPIE_REFERENCE", "/root/reference")


@pytest.fixture()
def files(tmp_path):
    rng = np.random.default_rng(5)
    out = {}
    for name, shape in (("colour", (37, 41, 3)), ("grey", (29, 33)), ("rgba", (31, 30, 4))):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        path = str(tmp_path / f"{name}.png")
        assert cv2.imwrite(path, img)
        out[name] = (path, img)
    return out


def want(img):
    """io.py:15-18 on what cv2.imread returns (which already drops alpha and expands grey to BGR)."""
    if img.ndim == 2:
        return np.stack([img] * 3, axis=-1)
    return img[..., :3]


@pytest.mark.parametrize("pinned", [True, False])
def test_read_image_contract(files, pinned):
    for name, (path, img) in files.items():
        got = bio.read_image(path, pinned=pinned)
        assert got.dtype == np.uint8 and got.ndim == 3 and got.shape[2] == 3 and got.flags["C_CONTIGUOUS"], name
        np.testing.assert_array_equal(got, want(img))
    with pytest.raises(FileNotFoundError):
        bio.read_image(files["colour"][0] + ".missing", pinned=pinned)


def test_read_images_default_mask_and_order(files, tmp_path):
    src_p, tgt_p = files["colour"][0], files["rgba"][0]
    with warnings.catch_warnings(record=True) as seen:
        warnings.simplefilter("always")
        src, mask, tgt = bio.read_images(src_p, str(tmp_path / "no-mask.png"), tgt_p)
    assert any("No mask file" in str(w.message) for w in seen)
    np.testing.assert_array_equal(src, want(files["colour"][1]))
    np.testing.assert_array_equal(tgt, want(files["rgba"][1]))
    assert mask.shape == src.shape and mask.dtype == np.uint8 and (mask == 255).all()
    src, mask, tgt = bio.read_images(src_p, files["grey"][0], tgt_p)
    np.testing.assert_array_equal(mask, want(files["grey"][1]))
    # three live results never share a staging buffer, and a released one is recycled
    assert not np.shares_memory(src, tgt) and not np.shares_memory(src, mask)
    with pytest.raises(FileNotFoundError):
        bio.read_images(src_p + ".missing", files["grey"][0], tgt_p)


def test_staging_buffers_are_recycled_only_when_released(files):
    a = bio.read_image(files["colour"][0])
    b = bio.read_image(files["colour"][0])
    assert not np.shares_memory(a, b)
    keep = a.copy()
    del a
    c = bio.read_image(files["colour"][0])  # may reuse a's buffer; b must be intact either way
    np.testing.assert_array_equal(b, keep)
    np.testing.assert_array_equal(c, keep)


def test_async_writer_snapshots_and_matches_imwrite(tmp_path):
    rng = np.random.default_rng(9)
    canvas = rng.integers(0, 256, (64, 48, 3), dtype=np.uint8)
    first = canvas.copy()
    bio.write_image_async(str(tmp_path / "a.png"), canvas)
    canvas[...] = 7  # the Processor reuses its canvas: the queued image must be the one handed in
    bio.write_image_async(str(tmp_path / "b.png"), canvas)
    bio.flush_writes()
    cv2.imwrite(str(tmp_path / "a_ref.png"), first)
    assert open(tmp_path / "a.png", "rb").read() == open(tmp_path / "a_ref.png", "rb").read()
    assert (cv2.imread(str(tmp_path / "b.png")) == 7).all()
    w = bio.ImageWriter()
    w.write(str(tmp_path / "no-such-dir" / "c.png"), canvas)  # cv2.imwrite fails quietly or raises: flush must not hang
    try:
        w.close()
    except Exception:
        pass


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "fpie")), reason="reference checkout not present")
def test_same_arrays_as_the_reference_io(files, tmp_path):
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_fpie_io", os.path.join(REF, "fpie", "io.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for name, (path, _) in files.items():
        np.testing.assert_array_equal(bio.read_image(path), ref.read_image(path), err_msg=name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = bio.read_images(files["colour"][0], str(tmp_path / "none.png"), files["grey"][0])
        exp = ref.read_images(files["colour"][0], str(tmp_path / "none.png"), files["grey"][0])
    for g, e in zip(got, exp, strict=True):
        np.testing.assert_array_equal(g, e)
        assert g.dtype == e.dtype
