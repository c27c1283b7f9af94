"""CPU tests of the counter-based random bits of the stochastic training mode: oracle/mode_rng.py (numpy) against the
engine's own header csrc/rng.cuh compiled for the host, plus the statistics the masks must have."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import mode_rng as R

ROOT = Path(__file__).resolve().parents[1]

HOST_PROG = r"""
#include <cstdio>
#define __host__
#define __device__
#define __forceinline__ inline
#include "rng.cuh"
int main() {
  using namespace mode;
  const unsigned long long seeds[3] = {0ull, 20261017ull, 0xfedcba9876543210ull};
  for (int s = 0; s < 3; ++s)
    for (unsigned step = 0; step < 3; ++step)
      for (unsigned stream = 1; stream <= 4; ++stream)
        for (unsigned layer = 0; layer < 3; ++layer) {
          const uint32_t key = rng_key(seeds[s], step, stream, layer);
          printf("%u", key);
          for (uint32_t idx = 0; idx < 4; ++idx) printf(" %u", rng_bits(key, idx * 2654435761u + 17u));
          printf("\n");
        }
  const float ps[5] = {0.0f, 0.1f, 0.3f, 0.5f, 0.999999f};
  for (int i = 0; i < 5; ++i) printf("T %u\n", drop_threshold(ps[i]));
  const uint32_t bs[4] = {0u, 511u, 0x80000000u, 0xffffffffu};
  for (int i = 0; i < 4; ++i) printf("U %.9g\n", rng_uniform(bs[i]));
  return 0;
}
"""


def test_numpy_bits_equal_the_engine_header(tmp_path):
    src = tmp_path / "rng_host.cpp"
    src.write_text(HOST_PROG)
    exe = tmp_path / "rng_host"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", str(ROOT / "mode_diffusion_policy_b200" / "csrc"), "-o", str(exe), str(src)],
                   check=True)
    lines = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    it = iter(lines)
    for seed in (0, 20261017, 0xFEDCBA9876543210):
        for step in range(3):
            for stream in range(1, 5):
                for layer in range(3):
                    want = [int(v) for v in next(it).split()]
                    key = R.rng_key(seed, step, stream, layer)
                    idx = (np.arange(4, dtype=np.uint64) * 2654435761 + 17) & 0xFFFFFFFF
                    assert [key] + R.rng_bits(key, idx).tolist() == want
    for p in (0.0, 0.1, 0.3, 0.5, 0.999999):
        assert next(it) == f"T {R.drop_threshold(p)}"
    for b in (0, 511, 0x80000000, 0xFFFFFFFF):
        assert np.float32(next(it).split()[1]) == np.float32(R.rng_uniform(np.uint32(b)))


@pytest.mark.parametrize("p", [0.1, 0.3])
def test_dropout_masks_hit_their_rate_and_are_decorrelated(p):
    a = R.attn_keep_mask(7, 0, 0, 64, 8, 14, p)
    m = R.mlp_keep_mask(7, 0, 1, np.arange(512), 2, 4, 1024, p)
    for keep in (a, m):
        assert abs(1.0 - keep.mean() - p) < 4e-3
    # different layer / step / seed / expert -> unrelated masks; same arguments -> the same mask
    assert np.array_equal(a, R.attn_keep_mask(7, 0, 0, 64, 8, 14, p))
    for other in (R.attn_keep_mask(7, 0, 1, 64, 8, 14, p), R.attn_keep_mask(7, 1, 0, 64, 8, 14, p),
                  R.attn_keep_mask(8, 0, 0, 64, 8, 14, p)):
        agree = (a == other).mean()
        assert abs(agree - ((1 - p) ** 2 + p ** 2)) < 1e-2
    m2 = R.mlp_keep_mask(7, 0, 1, np.arange(512), 3, 4, 1024, p)
    assert abs((m == m2).mean() - ((1 - p) ** 2 + p ** 2)) < 5e-3
    # neighbouring elements share a 32-bit word but use disjoint halves: no pair correlation
    both = (~m[:, 0::2] & ~m[:, 1::2]).mean()
    assert abs(both - p * p) < 3e-3
    g = R.goal_keep_mask(3, 5, 256, 512, p)
    assert abs(1.0 - g.mean() - p) < 4e-3


def test_multinomial_draws_follow_sampling_without_replacement():
    probs = np.array([[0.5, 0.25, 0.15, 0.1]], dtype=np.float32)
    T = 4000
    idx = R.multinomial_draws(11, 0, 0, probs, T, 2)
    assert (idx[:, 0] != idx[:, 1]).all()
    first = np.bincount(idx[:, 0], minlength=4) / T
    assert np.abs(first - probs[0]).max() < 0.03
    # P(second = j) = sum_i p_i p_j / (1 - p_i)
    p = probs[0].astype(np.float64)
    second = np.array([sum(p[i] * p[j] / (1 - p[i]) for i in range(4) if i != j) for j in range(4)])
    got = np.bincount(idx[:, 1], minlength=4) / T
    assert np.abs(got - second).max() < 0.03
    # a clamped, near one-hot row still yields K distinct experts
    hot = np.array([[1 - 3e-9, 1e-9, 1e-9, 1e-9]], dtype=np.float32)
    idx = R.multinomial_draws(1, 2, 3, hot, 64, 2)
    assert (idx[:, 0] == 0).all() and (idx[:, 1] != 0).all()
