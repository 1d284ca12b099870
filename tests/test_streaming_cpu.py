"""CPU: window bookkeeping of runtime/streaming.py with a fake engine (no kernels): windows start at 0, stride, 2*stride,
... for ragged pushes and for stride > seq_len (the old code advanced by fewer than `stride` frames)."""
import pytest
import torch

from runtime.streaming import StreamingVideoScorer


class _FakeEngine:
    def __init__(self):
        self.encoded = 0

    def encode(self, frames):
        self.encoded += frames.shape[0]
        return frames[:, :1, :1, :1].to(torch.bfloat16).permute(0, 2, 3, 1).contiguous(), None   # [n,1,1,1] "latent"

    def score_latents(self, lat, x, want_recon, want_heat):
        # the frames carry their stream index: return it so the test can see which frames a window covered
        assert lat.shape[:2] == x.shape[:2]
        assert torch.equal(lat[0, :, 0, 0, 0].float(), x[0, :, 0, 0, 0])
        return x[0, :, 0, 0, 0].clone()


class _FakeModel:
    def __init__(self):
        self.engine = _FakeEngine()

    def _get_engine(self, device):
        return self.engine


@pytest.mark.parametrize("n,T,stride,chunks", [
    (40, 16, 20, (7, 9, 3, 21)),      # stride > seq_len (the reported case: second window at 20, not 16)
    (50, 4, 11, (50,)),
    (50, 4, 11, (1,) * 50),
    (22, 8, 3, (5, 1, 13, 3)),
    (16, 16, 16, (16,)),
    (15, 16, 4, (15,)),               # never a full window
])
def test_window_starts(n, T, stride, chunks):
    assert sum(chunks) == n
    video = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1).expand(n, 3, 2, 2).contiguous()
    m = _FakeModel()
    sc = StreamingVideoScorer(m, seq_len=T, stride=stride)
    got, pos = [], 0
    for c in chunks:
        got += sc.push(video[pos:pos + c])
        pos += c
    starts = list(range(0, n - T + 1, stride))            # utils/video_dataset.py:371
    assert [s for s, _ in got] == starts
    for s, frames in got:
        assert frames.tolist() == list(range(s, s + T))
    assert m.engine.encoded <= n


def test_bad_arguments():
    with pytest.raises(ValueError):
        StreamingVideoScorer(_FakeModel(), seq_len=0)
    with pytest.raises(ValueError):
        StreamingVideoScorer(_FakeModel(), seq_len=4, stride=0)
    with pytest.raises(RuntimeError):
        StreamingVideoScorer(_FakeModel(), seq_len=4, stride=2).push(torch.zeros(3, 4, 8, 8))
