import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; without a device they are skipped, never silently passed.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def golden_head(compute_dtype=None):
    """The head whose weights generated tests/golden/head_b2p4.pt (incl. the calibrated score layer)."""
    import torch
    from ait_b200 import synth
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True,
                           compute_dtype=compute_dtype or torch.float32)
    g = load_golden("head_b2p4.pt")
    head.RCNN_cls_score.load_state_dict(g["cls_score_state"])
    return head, g


def head_inputs(B, P, first_unit=0):
    import torch
    from ait_b200 import synth
    non_img = torch.stack([synth.c4_map(first_unit + u) for u in range(B)])
    non_qry = torch.stack([synth.query_feat(first_unit + u) for u in range(B)])
    rois = torch.stack([synth.random_rois(first_unit + u, P, batch_index=u) for u in range(B)])
    return non_img, non_qry, rois
