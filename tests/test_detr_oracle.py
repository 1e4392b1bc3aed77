"""CPU: pins oracle/detr_oracle.py (the restatement of the DETR arithmetic) against golden vectors produced by
transformers' own DetrForObjectDetection + DetrImageProcessor (tests/golden/make_detr_golden.py) — the arithmetic the
reference's removed ViTDetector drove.  The reference's own tests hold no DETR vector."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import detr_oracle as do

from .conftest import GOLDEN


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN / "detr_small.npz"))


@pytest.fixture(scope="module")
def weights():
    return do.make_weights(0)


def test_preprocess_matches_transformers(golden):
    pv = do.preprocess(golden["small_frames"], do_resize=False).numpy()
    np.testing.assert_allclose(pv, golden["small_pixel_values"], rtol=0, atol=1e-6)
    # the 800/1333 rule + uint8 antialias bilinear resize + normalise: 180x320 -> 750x1333
    assert do.resized_size(180, 320) == (750, 1333) and do.resized_size(720, 1280) == (750, 1333)
    assert do.resized_size(800, 1333) == (800, 1333) and do.resized_size(1333, 800) == (1333, 800)
    pv = do.preprocess(golden["resized_frames"]).numpy()
    np.testing.assert_allclose(pv, golden["resized_pixel_values"], rtol=0, atol=1e-6)


def test_forward_fp32_matches_transformers(golden, weights):
    logits, boxes = do.forward(weights, golden["small_frames"], mode="fp32", do_resize=False)
    np.testing.assert_allclose(logits.numpy(), golden["small_logits"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(boxes.numpy(), golden["small_boxes"], rtol=1e-4, atol=2e-5)


def test_bf16_mode_tracks_fp32(golden, weights):
    """The bf16 rounding points move the outputs by bf16-level noise only (documents the H1 gap of SURVEY.md)."""
    l32, b32 = do.forward(weights, golden["small_frames"], mode="fp32", do_resize=False)
    l16, b16 = do.forward(weights, golden["small_frames"], mode="bf16", do_resize=False)
    assert float((l16 - l32).norm() / l32.norm()) < 5e-2
    assert float((b16 - b32).abs().max()) < 5e-2


def test_postprocess_and_detections(golden):
    logits, boxes = torch.from_numpy(golden["small_logits"]), torch.from_numpy(golden["small_boxes"])
    scores, labels, xyxy = do.postprocess(logits, boxes, 96, 128)
    assert scores.shape == (2, 100) and int(labels.max()) < 91
    thr = float(scores.flatten().median())
    dets = do.detections(logits, boxes, 96, 128, thr)
    n = sum(len(d) for d in dets)
    assert 0 < n < 200
    for rows in dets:
        for x, y, w, h, sc, fx, fy in rows:
            assert sc > thr and fx == pytest.approx(x + w / 2) and fy == pytest.approx(y + h)


def test_weights_are_non_degenerate(golden):
    """the seeded init must give distinct boxes per query and a usable person fraction."""
    boxes = golden["small_boxes"][0]
    assert np.unique(np.round(boxes, 3), axis=0).shape[0] > 90
    prob = torch.softmax(torch.from_numpy(golden["small_logits"]), -1)[..., :-1]
    frac_person = float((prob.argmax(-1) == do.PERSON_LABEL).float().mean())
    assert 0.2 < frac_person <= 1.0


@pytest.mark.parametrize("h0,w0,h1,w1", [(72, 128, 75, 133), (90, 160, 94, 167), (100, 90, 40, 37), (64, 64, 64, 200)])
def test_resize_restatement_is_bit_exact(h0, w0, h1, w1):
    """The numpy restatement of ATen's uint8 antialias kernel (which the CUDA resize is checked against on the GPU)."""
    img = np.random.default_rng(h0 * w1).integers(0, 256, (2, 3, h0, w0), dtype=np.uint8)
    ref = torch.nn.functional.interpolate(torch.from_numpy(img), size=(h1, w1), mode="bilinear", antialias=True,
                                          align_corners=False).numpy()
    assert (do.resize_u8_antialias(img, h1, w1) == ref).all()


def test_roi_features_match_the_reference_feature_extractor():
    """oracle.roi_features == the reference's own FeatureExtractor.extract_roi_features output
    (tests/golden/make_roi_golden.py imports src/tracking/feature_extractor.py): inside / clipped / degenerate boxes."""
    g = np.load(GOLDEN / "roi_golden.npz")
    got = do.roi_features(g["feat"], [tuple(b) for b in g["boxes"]], tuple(int(v) for v in g["image_shape"]))
    assert got.dtype == g["features"].dtype and got.shape == g["features"].shape
    np.testing.assert_array_equal(got, g["features"])
    assert do.roi_features(g["feat"], [], (720, 1280)).shape == (0, 64)


@pytest.mark.parametrize("h0,w0", [(800, 1333), (720, 1280)])
def test_oracle_matches_transformers_full_size(weights, h0, w0):
    """The pin at BASELINE's frame sizes (the golden vectors are 96x128 only): oracle fp32 mode against transformers'
    own DetrImageProcessor + DetrForObjectDetection, live, on one synthetic frame of the config-2 / config-1 size - preprocessing
    (uint8 antialias resize for 720p), logits and boxes."""
    import os

    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    from transformers import DetrImageProcessor

    frames = do.synthetic_frames(1, h0, w0, seed=17)
    model = do.hf_model(weights)
    inp = DetrImageProcessor()(images=[np.ascontiguousarray(frames[0][:, :, ::-1])], return_tensors="pt")
    with torch.no_grad():
        out = model(**inp)
    pv = do.preprocess(frames)
    assert tuple(pv.shape) == tuple(inp["pixel_values"].shape)
    np.testing.assert_allclose(pv.numpy(), inp["pixel_values"].numpy(), rtol=0, atol=1e-6)
    logits, boxes = do.forward(weights, frames, mode="fp32")
    np.testing.assert_allclose(logits.numpy(), out.logits.numpy(), rtol=1e-4, atol=5e-4)
    np.testing.assert_allclose(boxes.numpy(), out.pred_boxes.numpy(), rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("sizes,do_resize", [([(96, 128), (80, 144), (64, 64)], False), ([(720, 1280), (800, 1333)], True)])
def test_oracle_mixed_batch_matches_transformers(weights, sizes, do_resize):
    """Batches that mix frame sizes: DetrImageProcessor pads to the batch maximum and returns pixel_mask, DetrForObjectDetection
    carries the mask through the sine embedding and the attentions.  oracle.forward_mixed (fp32) against transformers, live: small
    un-resized frames of three sizes, and a 720p + 800x1333 pair through the 800 / 1333 resize rule."""
    import os

    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    from transformers import DetrImageProcessor

    frames = [do.synthetic_frames(1, h, w, seed=30 + i)[0] for i, (h, w) in enumerate(sizes)]
    model = do.hf_model(weights)
    inp = DetrImageProcessor(do_resize=do_resize)(images=[np.ascontiguousarray(f[:, :, ::-1]) for f in frames], return_tensors="pt")
    assert "pixel_mask" in inp and not bool(inp["pixel_mask"].all())        # the batch really is padded
    with torch.no_grad():
        out = model(**inp)
    taps: dict = {}
    logits, boxes = do.forward_mixed(weights, frames, mode="fp32", taps=taps, do_resize=do_resize)
    np.testing.assert_allclose(taps["pixel_values"].numpy(), inp["pixel_values"].numpy(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(logits.numpy(), out.logits.numpy(), rtol=1e-4, atol=5e-4)
    np.testing.assert_allclose(boxes.numpy(), out.pred_boxes.numpy(), rtol=1e-4, atol=5e-5)
