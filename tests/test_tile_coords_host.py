"""CPU: the GEMM's tile rasterisation (easynlp_b200/csrc/gemm_tiles.h, the same source the kernel compiles) visits every output
tile exactly once -- plain and band order, ragged band at the bottom -- and the band order really walks down a band first."""
import os
import subprocess

SRC = r'''
#include <cstdio>
#include <vector>
#include "gemm_tiles.h"
int main() {
  const int shapes[][3] = {{1,1,1},{394,9,1},{394,9,16},{7,3,16},{16,65,16},{17,65,16},{33,100,16},{2048,1024,16},{5,70,4},{31,67,8}};
  for (auto& sh : shapes) {
    const int mt = sh[0], nt = sh[1], gm = sh[2];
    std::vector<int> seen(mt * nt, 0);
    for (int t = 0; t < mt * nt; ++t) {
      int mi = -1, ni = -1;
      clipk::tile_coords_raw(mt, nt, gm, t, mi, ni);
      if (mi < 0 || mi >= mt || ni < 0 || ni >= nt) { std::printf("FAIL range %d %d %d t=%d -> %d %d\n", mt, nt, gm, t, mi, ni); return 1; }
      seen[mi * nt + ni]++;
    }
    for (int v : seen) if (v != 1) { std::printf("FAIL bijection %d %d %d\n", mt, nt, gm); return 1; }
    if (gm > 1 && mt >= gm) {   // the first gm ids share column tile 0 and walk rows 0..gm-1
      for (int t = 0; t < gm; ++t) { int mi, ni; clipk::tile_coords_raw(mt, nt, gm, t, mi, ni); if (mi != t || ni != 0) { std::printf("FAIL order\n"); return 1; } }
    }
  }
  std::printf("OK\n");
  return 0;
}
'''


def test_tile_coords_bijection(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.cpp"; src.write_text(SRC)
    exe = str(tmp_path / "t")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(root, "easynlp_b200", "csrc"), str(src), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr
