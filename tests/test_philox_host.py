"""CPU: the dropout RNG (philox4x32 in easynlp_b200/csrc/common.cuh, compiled for the host by nvcc) against the published
Philox4x32-10 known-answer vectors (Random123 kat_vectors: zero / all-ones / pi-digits counters and keys)."""
import os
import shutil
import subprocess

import pytest

SRC = r'''
#include <cstdio>
#include "common.cuh"
int main() {
  struct KAT { uint32_t c[4], k[2], want[4]; };
  const KAT kats[] = {
    {{0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}, {0x00000000u, 0x00000000u}, {0x6627e8d5u, 0xe169c58du, 0xbc57ac4cu, 0x9b00dbd8u}},
    {{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}, {0x408f276du, 0x41c83b0eu, 0xa20bc7c6u, 0x6d5451fdu}},
    {{0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u}, {0xa4093822u, 0x299f31d0u}, {0xd16cfe09u, 0x94fdccebu, 0x5001e420u, 0x24126ea1u}},
  };
  for (const KAT& t : kats) {
    const uint4 r = clipk::philox4x32(t.c[0], t.c[1], t.c[2], t.c[3], t.k[0], t.k[1]);
    if (r.x != t.want[0] || r.y != t.want[1] || r.z != t.want[2] || r.w != t.want[3]) {
      std::printf("FAIL %08x %08x %08x %08x\n", r.x, r.y, r.z, r.w);
      return 1;
    }
  }
  std::printf("OK\n");
  return 0;
}
'''


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc (host compile only, no GPU)")
def test_philox4x32_10_known_answers(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = tmp_path / "kat.cu"; src.write_text(SRC)
    exe = str(tmp_path / "kat")
    subprocess.check_call([nvcc, "-std=c++17", "-arch=sm_100a", "-I", os.path.join(root, "easynlp_b200", "csrc"), str(src), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr
