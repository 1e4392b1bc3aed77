"""GPU: the Phase 2 -> 3 -> count flow end to end, two ways that must agree:
  (1) the reference-shaped objects a Phase would drive (src/pipeline/phases/detection.py:91-103, transform.py:279-308,
      aggregation.py:38-43): ViTDetector.detect_batch -> HomographyTransformer.transform_batch -> ZoneClassifier.classify ->
      Aggregator.get_zone_counts, one Python object per detection;
  (2) the tensor pipeline (DetectCountPipeline.run_tensors) + the tensor-side exporter."""

from __future__ import annotations

import json

import numpy as np
import pytest

from oracle import detr_oracle as do
from oracle import floor_oracle as fo

pytestmark = pytest.mark.gpu


def test_object_flow_equals_tensor_flow(built_lib, tmp_path):
    import torch

    from office_person_detection_vit_b200.aggregation import Aggregator
    from office_person_detection_vit_b200.detection import ViTDetector
    from office_person_detection_vit_b200.export import export_results, frame_results_from_tensors
    from office_person_detection_vit_b200.pipeline import DetectCountPipeline
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    torch.cuda.init()
    zones = fo.star_zones(16, seed=2)
    zone_ids = [z["id"] for z in zones]
    det = ViTDetector(confidence_threshold=0.3, state_dict=do.make_weights(0))
    det.load_model()
    det.model.set_resize(False)          # small frames fed as they are (DetrImageProcessor(do_resize=False))
    tr = HomographyTransformer(fo.H_CONFIG, FloorMapConfig())
    zc = ZoneClassifier(zones, allow_overlap=False)
    frames = do.synthetic_frames(3, 192, 256, seed=9)
    timestamps = [f"2025/08/26 16:0{i}:00" for i in range(3)]

    # (1) object flow
    agg = Aggregator()
    per_frame = det.detect_batch(list(frames))
    counts_obj = []
    for ts, dets in zip(timestamps, per_frame):
        for d, res in zip(dets, tr.transform_batch([d.bbox for d in dets])):
            d.floor_coords, d.floor_coords_mm = res.floor_coords_px, res.floor_coords_mm
            d.zone_ids = zc.classify(res.floor_coords_px)
        counts_obj.append(agg.aggregate_frame(ts, dets))
    assert sum(len(d) for d in per_frame) > 10

    # (2) tensor flow
    pipe = DetectCountPipeline(det, tr, zc)
    out = pipe.run_tensors(torch.from_numpy(frames).cuda())
    torch.cuda.synchronize()
    assert out["n_keep"].cpu().tolist() == [len(d) for d in per_frame]
    counts_tensor = zc.counts_to_dicts(out["hist"])
    assert [dict(sorted(c.items())) for c in counts_tensor] == [dict(sorted(c.items())) for c in counts_obj]

    # records and JSON from the tensors == records from the object flow
    frs = frame_results_from_tensors(out, [10, 11, 12], timestamps, tr, zone_ids)
    for fr, dets in zip(frs, per_frame):
        assert len(fr.detections) == len(dets)
        for a, b in zip(fr.detections, dets):
            assert a.bbox == pytest.approx(b.bbox) and a.zone_ids == b.zone_ids
            assert a.floor_coords == pytest.approx(b.floor_coords, rel=1e-12, abs=1e-9)
            assert a.floor_coords_mm == pytest.approx(b.floor_coords_mm, rel=1e-12, abs=1e-7)
    path = export_results(out, [10, 11, 12], timestamps, tr, zone_ids, tmp_path)
    data = json.loads(path.read_text(encoding="utf-8"))
    assert data["transform_method"] == "homography" and len(data["frames"]) == 3
    d0 = data["frames"][0]["detections"][0]
    assert set(d0) >= {"bbox", "confidence", "camera_coords", "floor_coords_px", "floor_coords_mm"}
    assert d0["bbox"]["x"] == round(per_frame[0][0].bbox[0], 6)

    # CSV through the aggregator's dense-histogram entry
    agg2 = Aggregator()
    agg2.aggregate_histogram(timestamps, out["hist"], zone_ids)
    agg.export_csv(str(tmp_path / "a.csv"), zone_ids)
    agg2.export_csv(str(tmp_path / "b.csv"), zone_ids)
    assert (tmp_path / "a.csv").read_text() == (tmp_path / "b.csv").read_text()


def test_piecewise_affine_pipeline(built_lib):
    """The reference's default transform.method on the tensor path: detector -> PiecewiseAffineTransformer -> zones -> counts,
    against the object flow (transform_batch + classify + get_zone_counts on Detection records)."""
    import torch

    from office_person_detection_vit_b200.aggregation import Aggregator
    from office_person_detection_vit_b200.detection import ViTDetector
    from office_person_detection_vit_b200.pipeline import DetectCountPipeline
    from office_person_detection_vit_b200.transform import FloorMapConfig, PiecewiseAffineTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    from .conftest import GOLDEN

    g = np.load(GOLDEN / "pwa_golden.npz")
    # correspondences scaled from the 1280x720 camera frame of the golden set to the 256x192 test frames
    src = g["src"] * np.array([256 / 1280, 192 / 720])
    tr = PiecewiseAffineTransformer(src, g["dst"], FloorMapConfig())
    zones = fo.grid_zones(16)
    zc = ZoneClassifier(zones, allow_overlap=False)
    det = ViTDetector(confidence_threshold=0.3, state_dict=do.make_weights(0))
    det.load_model()
    det.model.set_resize(False)
    frames = do.synthetic_frames(2, 192, 256, seed=9)
    out = DetectCountPipeline(det, tr, zc).run_tensors(torch.from_numpy(frames).cuda())
    torch.cuda.synchronize()
    agg = Aggregator()
    per_frame = det.detect_batch(list(frames))
    assert out["n_keep"].cpu().tolist() == [len(d) for d in per_frame] and sum(len(d) for d in per_frame) > 10
    counts = []
    for dets in per_frame:
        for d, res in zip(dets, tr.transform_batch([d.bbox for d in dets])):
            d.zone_ids = zc.classify(res.floor_coords_px)
        counts.append(agg.get_zone_counts(dets))
    got = zc.counts_to_dicts(out["hist"])
    assert [dict(sorted(c.items())) for c in got] == [dict(sorted(c.items())) for c in counts]


def test_run_stream_equals_run_tensors(built_lib):
    """Double-buffered ingest (pinned staging + copy stream) yields exactly what run_tensors yields on the same batches, in
    order, with each batch's counts in its own timestamp rows."""
    import torch

    from office_person_detection_vit_b200.detection import ViTDetector
    from office_person_detection_vit_b200.pipeline import DetectCountPipeline
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    det = ViTDetector(confidence_threshold=0.3, state_dict=do.make_weights(0))
    det.load_model()
    det.model.set_resize(False)
    pipe = DetectCountPipeline(det, HomographyTransformer(fo.H_CONFIG, FloorMapConfig()), ZoneClassifier(fo.grid_zones(16), allow_overlap=False))
    batches = [do.synthetic_frames(2, 160, 224, seed=40 + i) for i in range(5)]
    hist = torch.zeros(10, 17, dtype=torch.int32, device="cuda")
    got = [(o["n_keep"].clone(), o["det_xywh"].clone(), o["zone_idx"].clone()) for o in pipe.run_stream(iter(batches), hist=hist)]
    torch.cuda.synchronize()
    assert len(got) == 5
    ref_hist = torch.zeros(10, 17, dtype=torch.int32, device="cuda")
    for i, b in enumerate(batches):
        o = pipe.run_tensors(torch.from_numpy(b).cuda(), hist=ref_hist, slot_base=2 * i)
        assert torch.equal(got[i][0], o["n_keep"]) and torch.equal(got[i][1], o["det_xywh"]) and torch.equal(got[i][2], o["zone_idx"])
    assert torch.equal(hist, ref_hist) and int(hist.sum()) == int(sum(int(g[0].sum()) for g in got))


def test_captured_step_equals_run_tensors(built_lib):
    """DetectCountPipeline.capture: replaying the CUDA graph on refilled input buffers gives what run_tensors gives."""
    import torch

    from office_person_detection_vit_b200.detection import ViTDetector
    from office_person_detection_vit_b200.pipeline import DetectCountPipeline
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    det = ViTDetector(confidence_threshold=0.3, state_dict=do.make_weights(0))
    det.load_model()
    det.model.set_resize(False)
    pipe = DetectCountPipeline(det, HomographyTransformer(fo.H_CONFIG, FloorMapConfig()), ZoneClassifier(fo.grid_zones(16), allow_overlap=False))
    buf = torch.from_numpy(do.synthetic_frames(2, 160, 224, seed=50)).cuda()
    hist = torch.zeros(2, 17, dtype=torch.int32, device="cuda")
    replay = pipe.capture(buf, hist=hist, zero_hist=True)
    for seed in (51, 52, 53):
        frames = torch.from_numpy(do.synthetic_frames(2, 160, 224, seed=seed)).cuda()
        buf.copy_(frames)
        out = replay()
        got = (out["n_keep"].clone(), out["det_xywh"].clone(), out["zone_idx"].clone(), hist.clone())
        ref_hist = torch.zeros(2, 17, dtype=torch.int32, device="cuda")
        ref = pipe.run_tensors(frames, hist=ref_hist)
        torch.cuda.synchronize()
        assert torch.equal(got[0], ref["n_keep"]) and torch.equal(got[1], ref["det_xywh"]) and torch.equal(got[2], ref["zone_idx"])
        assert torch.equal(got[3], ref_hist) and int(ref_hist.sum()) > 0
