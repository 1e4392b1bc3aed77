"""CPU: the tensor-side exporter (office_person_detection_vit_b200/export) writes byte-identical
coordinate_transformations.json to the reference's TransformPhase.export_results (src/pipeline/phases/transform.py:398-531);
golden texts produced by the reference itself (tests/golden/make_export_golden.py)."""

from __future__ import annotations

import json

import numpy as np
import pytest

from office_person_detection_vit_b200.export import dumps_coordinate_transformations, format_coordinate_transformations

from .conftest import GOLDEN


@pytest.fixture(scope="module")
def golden():
    return json.loads((GOLDEN / "export_golden.json").read_text(encoding="utf-8"))


@pytest.mark.parametrize("case", [0, 1, 2])
def test_json_is_byte_identical_to_the_reference(golden, case):
    i = golden["inputs"]
    c = golden["cases"][case]
    data = format_coordinate_transformations(
        np.array(i["n_keep"]), np.array(i["xywh"]), np.array(i["score"]), np.array(i["foot"]), np.array(i["px"]),
        np.array(i["mm"]), np.array(i["zone_idx"]), i["frame_numbers"], i["timestamps"], i["transformer_info"], i["zone_ids"],
        "homography", c["json_optimization"])
    assert dumps_coordinate_transformations(data, c["json_optimization"]) == c["text"]
