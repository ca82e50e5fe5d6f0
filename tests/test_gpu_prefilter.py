"""GPU parity: the greedy frame filters around the frame-pair kernel:
key-frame pre-filter decisions (bp:179-228) and QA re-decode de-duplication (hm:2226-2249, hm:2789-2812)."""
import numpy as np
import pytest

import cases
from oracle import hippo_oracle as O
from test_oracle import _decoded_prefilter_frames

pytestmark = pytest.mark.gpu


def test_prefilter_matches_reference_outputs(cuda_device, tmp_path):
    """Same MJPG stream, same decoder: the frames the reference's extract_frames_from_video saved."""
    from hippomm_b200.prefilter import select_saved_frames

    decoded = _decoded_prefilter_frames(tmp_path)
    g = cases.golden()
    numbers, times = select_saved_frames(decoded, **cases.PREFILTER_PARAMS)
    assert numbers == g["prefilter_frame_numbers"].tolist()
    assert times == g["prefilter_frame_times"].tolist()


@pytest.mark.parametrize("window", [1, 3, 8])
@pytest.mark.parametrize("params", [dict(video_fps=30.0, max_diff_threshold=0.3, check_interval=10),
                                    dict(video_fps=30.0, max_diff_threshold=0.3, check_interval=30),
                                    dict(video_fps=24.0, max_diff_threshold=0.05, check_interval=7),
                                    dict(video_fps=10.0, max_diff_threshold=0.9, check_interval=1)])
def test_prefilter_matches_oracle_on_raw_frames(cuda_device, params, window):
    """The speculation window must not change the decisions; thresholds / intervals / rates vary the chain."""
    from hippomm_b200.prefilter import select_saved_frames

    frames = cases.prefilter_frames()
    want = O.select_saved_frames(frames, **params)
    # band 0: fallback launches only; 2: the table misses whenever the anchor is more than two candidates back; 24: table
    for band in (0, 2, 24):
        got = select_saved_frames(frames, window=window, band=band, **params)
        assert got[0] == want[0] and got[1] == want[1], band


def test_prefilter_edge_cases(cuda_device):
    from hippomm_b200.prefilter import select_saved_frames

    frames = cases.prefilter_frames()
    assert select_saved_frames(frames[:0], 30.0) == ([], [])
    assert select_saved_frames(frames[:1], 30.0) == ([0], [0.0])
    tiny = frames[:, :5, :5]                                            # smaller than the SSIM window: MSE fallback (bp:64-71)
    want = O.select_saved_frames(tiny, 30.0, 0.001, 5)
    assert select_saved_frames(tiny, 30.0, 0.001, 5) == want
    const = np.full((70, 32, 32, 3), 7, dtype=np.uint8)                 # SSIM NaN -> MSE fallback = 0 -> only frame 0
    assert select_saved_frames(const, 30.0, 0.3, 10) == ([0], [0.0])


@pytest.mark.parametrize("threshold", [0.3, 0.4])
def test_dedup_matches_oracle(cuda_device, threshold):
    from hippomm_b200.prefilter import dedup_window_frames

    for i, win in enumerate(cases.dedup_windows()):
        for window in (1, 4):
            for band in (0, 3, 128):
                assert dedup_window_frames(win, threshold, window=window, band=band) == \
                    O.dedup_window_frames(win, threshold), (i, window, band)
