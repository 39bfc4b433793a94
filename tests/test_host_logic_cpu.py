"""Host-side planning logic that needs no GPU: the launch-list recorder (incl. the per-step time-embedding reuse of the
fused PC loop), the per-tap kernel's pixel-box choice and the transposed kernel's macro-tile rule (which csrc/conv_gemm.cu
mirrors in pick_t_rows)."""
import math

import pytest

from conditional_score_diffusion_b200 import engine as E
from conditional_score_diffusion_b200 import kernels as K


def test_recorder_runs_in_order_and_skips_by_function():
    log = []

    def a(x):
        log.append(("a", x))

    def b(x, y=0):
        log.append(("b", x, y))

    rec = E.Recorder()
    rec.add(a, 1)
    rec.add(b, 2, y=3)
    rec.add(a, 4)
    rec.run()
    assert log == [("a", 1), ("b", 2, 3), ("a", 4)]
    log.clear()
    rec.run(skip=(b,))                       # what Plan.launch(reuse_time_embedding=True) does with the two temb launches
    assert log == [("a", 1), ("a", 4)]
    assert len(rec) == 3


@pytest.mark.parametrize("h,w,batch", [(160, 160, 64), (20, 20, 64), (10, 10, 64), (5, 5, 64), (5, 5, 3), (16, 16, 2),
                                       (1, 400, 64), (13, 7, 5)])
def test_pick_tile_covers_the_tensor_with_at_most_128_rows(h, w, batch):
    tw, th, tb = K.pick_tile(h, w, batch)
    assert 1 <= tw <= w and 1 <= th <= h and 1 <= tb <= batch
    assert tw * th * tb <= 128, "one CTA tile = the 128 rows of an M = 128 MMA"
    tiles = math.ceil(w / tw) * math.ceil(h / th) * math.ceil(batch / tb)
    assert tiles * 128 >= h * w * batch
    # never worse than the trivial row tiling
    trivial = math.ceil(w / min(w, 128)) * h * batch
    assert tiles <= trivial


def test_transposed_tile_rows_rule():
    # 160 -> 5 exact 32-row tiles, 80 -> 28 (3 tiles, 5 % padding instead of 17 %), 40 -> 2 exact 20-row tiles
    assert K.transposed_tile_rows(160) == 32
    assert K.transposed_tile_rows(80) == 28
    assert K.transposed_tile_rows(40) == 20
    for h in (20, 32, 40, 64, 80, 128, 160, 256):
        t = K.transposed_tile_rows(h)
        assert t in (20, 24, 28, 32) and math.ceil(h / t) * t >= h


def test_transposed_shape_rule():
    assert K.transposed_shape_ok(160, 160) and K.transposed_shape_ok(80, 80) and K.transposed_shape_ok(40, 40)
    assert not K.transposed_shape_ok(20, 20), "20 px levels stay on the per-tap kernel (ragged tiles measured slower)"
    assert not K.transposed_shape_ok(10, 10) and not K.transposed_shape_ok(5, 5)
