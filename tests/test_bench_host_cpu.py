"""bench.py host logic without a GPU: the synthetic batch contract of SURVEY.md §8d, the algorithmic-FLOP table, the dominant-launch /
traffic bookkeeping, and the JSON line of the reference arm (`--impl reference`, the oracle port on the host cores)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_synth_batch_contract():
    image, ids = bench.synth_batch(64, 32, 77, 21128, 1234)
    image2, ids2 = bench.synth_batch(64, 32, 77, 21128, 1234)
    assert torch.equal(image, image2) and torch.equal(ids, ids2)  # seeded
    assert image.shape == (64, 3, 32, 32) and ids.shape == (64, 77) and ids.dtype == torch.int64
    assert (ids[:, 0] == 101).all()
    lens = (ids != 0).sum(1)
    assert int(lens.min()) >= 8 and int(lens.max()) <= 77 and len(set(lens.tolist())) > 5  # ragged: the key-mask path is exercised
    for row, n in zip(ids, lens.tolist()):
        assert int(row[n - 1]) == 102 and (row[n:] == 0).all() and (row[:n] != 0).all()


def test_flop_table_matches_survey_formula():
    vit_l = (24 * 1024**2 + 4 * 257 * 1024) * 24 * 257 + 2 * 256 * 588 * 1024
    bert_b = (24 * 768**2 + 4 * 77 * 768) * 12 * 77
    assert abs(bench.FLOP_PER_PAIR["ViT-L-14"] - 3 * (vit_l + bert_b)) < 0.01 * bench.FLOP_PER_PAIR["ViT-L-14"]
    assert bench.FLOP_PER_PAIR["M2-Encoder-1B"] > bench.FLOP_PER_PAIR["M2-Encoder-0.4B"] > 1e11


class _Ev:
    def __init__(self, t):
        self.t = t

    def elapsed_time(self, other):
        return other.t - self.t


def test_dominant_launch_picks_largest_share_and_quotes_traffic():
    # (flops, start, stop, splits, (M, N, K, a_mn, b_mn, bias, act, aux, dact, res, f32))
    fc = (263168, 4096, 1024, 0, 0, 1, 1, 1, 0, 0, 0)
    pj = (263168, 1024, 4096, 0, 0, 1, 0, 0, 0, 1, 0)
    prof = [(2.0 * 263168 * 4096 * 1024, _Ev(0.0), _Ev(2.1), 1, fc)] * 3 + [(2.0 * 263168 * 1024 * 4096, _Ev(0.0), _Ev(1.8), 1, pj)] * 3
    d = bench.dominant_launch(prof, 1407.1)
    assert d["kernel"] == "gemm_tcgen05_kernel<0, 0, 0, 2, 67>" and d["shape_MNK"] == [263168, 4096, 1024] and d["launches"] == 3
    assert abs(d["achieved"] - 2.0 * 263168 * 4096 * 1024 / 2.1e-3 / 1e12) < 0.1
    assert d["algorithmic_bytes"] == 2 * (263168 * 1024 + 4096 * 1024) + 2 * 263168 * 4096 * 2 + 2 * 4096
    committed = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_latest.json")))
    assert d["traffic"] == committed["kernels"][d["kernel"]]["dram_bytes"]
    assert 0.9 < d["traffic"] / d["algorithmic_bytes"] < 1.3  # measured DRAM bytes ~ algorithmic: no wasted re-reads
    assert bench.dominant_launch([], 1407.1) is None


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "1",
                        "--model", "ViT-B-16"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    # the unmodified reference modules when baseline/_ref is installed (build() does that wherever /root/reference exists), else the port
    expect = "reference" if os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "antmmf/modules/vision/backbone/clip/cn_model.py")) else "port"
    assert d["cpu_baseline"]["kind"] == expect and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
