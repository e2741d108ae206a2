"""The optional viewer (fish_abm_b200/viewer.py) renders on the host from plain arrays: testable without a GPU."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from fish_abm_b200 import synth, viewer  # noqa: E402


def test_environment_painting_matches_the_reference_rule(q100):
    img = viewer.environment_image(q100, 2000, 2000)
    assert img.shape == (2000, 2000, 3)
    # fish_mod.cpp:634: Vec3b(q*(255/2), ...) with integer division; cells are 20x20 pixels, [y/20][x/20]
    for (r, c) in [(0, 0), (17, 63), (99, 99)]:
        assert np.all(img[r * 20:(r + 1) * 20, c * 20:(c + 1) * 20] == min(255, int(q100[r, c]) * 127))


def test_frame_shows_fish_over_a_dim_environment(q100):
    st = synth.lattice_state(60, 2000, 2000, seed=1)
    f = viewer.draw_frame(st, q100)
    assert f.shape == (2000, 2000, 3) and f.dtype == np.uint8
    env_only = viewer.draw_frame({k: v[:0] for k, v in st.items()}, q100)
    assert env_only.max() <= 26            # 0.1 * 254, the addWeighted blend (fish_mod.cpp:144)
    x, y = (int(v) for v in st["coords"][0])
    assert f[y - 12:y + 12, x - 12:x + 12].max() > env_only[y - 12:y + 12, x - 12:x + 12].max()  # an arrow is there


@pytest.mark.gpu
def test_record_steps_the_school_and_writes_a_video(q100, tmp_path):
    from fish_abm_b200 import School
    with School(n=60, seed=4) as S:
        S.initialize(5.0)
        S.set_quality(q100)
        frames = viewer.record(S, 20, every=5, path=str(tmp_path / "fish.avi"))
        assert len(frames) == 4 and S.step_index == 20
    assert (tmp_path / "fish.avi").stat().st_size > 1000
