// TEST INFRASTRUCTURE.  Compiles easy_gaussian_splatting_b200/csrc/egs_math.cuh with the HOST
// compiler (g++ -ffp-contract=off) and loops it over Gaussians on the CPU, mirroring what
// projection_fwd_kernel / projection_bwd_kernel do per thread.  This lets the CPU test suite check
// the hand-derived projection / SH formulas (and their bit-exactness against the oracle) without
// a GPU.  It is never loaded by the product package.
#include <cstdint>
#include <cstring>
#include <cmath>
using std::isfinite;
#include "../../easy_gaussian_splatting_b200/csrc/egs_math.cuh"

using namespace egs;

extern "C" void hh_projection_fwd(int C, int N, const float* means, const float* quats, const float* scales,
                                  const float* sh, int K, int deg, const float* viewmats, const float* Ks, int W,
                                  int H, float eps2d, float near_plane, float far_plane, float radius_clip,
                                  int tile_size, int tw, int th, int32_t* radii, float* means2d, float* depths,
                                  float* conics, float* colors, int32_t* tiles, float* comps /* nullable */) {
  for (int c = 0; c < C; ++c) {
    Camera cam;
    load_camera(viewmats + c * 16, Ks + c * 9, cam);
    for (int n = 0; n < N; ++n) {
      size_t idx = (size_t)c * N + n;
      ProjState st;
      ProjOut o;
      bool vis = project_fwd(means + 3 * n, quats + 4 * n, scales + 3 * n, cam, (float)W, (float)H, eps2d, near_plane,
                             far_plane, radius_clip, st, o);
      int nt = 0;
      float rgb[3] = {0, 0, 0};
      if (vis) {
        int x0, y0, x1, y1;
        tile_rect(o.m2x, o.m2y, o.radius, (float)tile_size, tw, th, x0, y0, x1, y1);
        nt = (x1 - x0) * (y1 - y0);
        if (deg >= 0) {
          float co[48];
          int nb = (deg + 1) * (deg + 1);
          for (int i = 0; i < 48; ++i) co[i] = i < nb * 3 ? sh[(size_t)n * K * 3 + i] : 0.f;
          sh_color_fwd(deg, means + 3 * n, cam, co, rgb);
        }
      }
      radii[idx] = o.radius; tiles[idx] = nt;
      means2d[idx * 2] = o.m2x; means2d[idx * 2 + 1] = o.m2y; depths[idx] = o.depth;
      conics[idx * 3] = o.ca; conics[idx * 3 + 1] = o.cb; conics[idx * 3 + 2] = o.cc;
      colors[idx * 3] = rgb[0]; colors[idx * 3 + 1] = rgb[1]; colors[idx * 3 + 2] = rgb[2];
      if (comps) comps[idx] = o.comp;
    }
  }
}

extern "C" void hh_projection_bwd(int C, int N, const float* means, const float* quats, const float* scales,
                                  const float* sh, int K, int deg, const float* viewmats, const float* Ks, int W,
                                  int H, float eps2d, const int32_t* radii, const float* colors,
                                  const float* v_means2d, const float* v_conics, const float* v_colors,
                                  float* v_means, float* v_quats, float* v_scales, float* v_sh,
                                  const float* v_comps /* nullable: antialiased-mode compensation gradients */) {
  memset(v_means, 0, sizeof(float) * 3 * N);
  memset(v_quats, 0, sizeof(float) * 4 * N);
  memset(v_scales, 0, sizeof(float) * 3 * N);
  memset(v_sh, 0, sizeof(float) * (size_t)N * K * 3);
  for (int c = 0; c < C; ++c) {
    Camera cam;
    load_camera(viewmats + c * 16, Ks + c * 9, cam);
    for (int n = 0; n < N; ++n) {
      size_t idx = (size_t)c * N + n;
      if (radii[idx] <= 0) continue;
      ProjState st;
      ProjOut o;
      project_fwd(means + 3 * n, quats + 4 * n, scales + 3 * n, cam, (float)W, (float)H, eps2d, 0.f, INFINITY, -1.f, st, o);
      if (o.radius > 0)
        project_bwd(st, scales + 3 * n, cam, v_means2d[idx * 2], v_means2d[idx * 2 + 1], 0.f, v_conics[idx * 3],
                    v_conics[idx * 3 + 1], v_conics[idx * 3 + 2], o, v_means + 3 * n, v_quats + 4 * n, v_scales + 3 * n,
                    v_comps ? v_comps[idx] : 0.f, eps2d);
      if (deg >= 0) {
        float co[48], vco[48];
        int nb = (deg + 1) * (deg + 1);
        for (int i = 0; i < 48; ++i) { co[i] = i < nb * 3 ? sh[(size_t)n * K * 3 + i] : 0.f; vco[i] = 0.f; }
        sh_color_bwd(deg, means + 3 * n, cam, co, colors + idx * 3, v_colors + idx * 3, vco, v_means + 3 * n);
        for (int i = 0; i < nb * 3; ++i) v_sh[(size_t)n * K * 3 + i] += vco[i];
      }
    }
  }
}

// Tight tile rectangles exactly as the projection kernels produce them (tile_rect -> sigma_cutoff ->
// tighten_tile_rect -> pack_tile_rect), from 2-D means / radii / conics / opacities.
extern "C" void hh_tight_rects(int n, const float* means2d, const int32_t* radii, const float* conics, const float* opacities,
                               int tile_size, int tw, int th, int32_t* classic /*[n,4] x0 y0 x1 y1*/, int32_t* packed /*[n,2]*/) {
  for (int i = 0; i < n; ++i) {
    int32_t x0, y0, x1, y1;
    tile_rect(means2d[2 * i], means2d[2 * i + 1], radii[i], (float)tile_size, tw, th, x0, y0, x1, y1);
    classic[4 * i] = x0; classic[4 * i + 1] = y0; classic[4 * i + 2] = x1; classic[4 * i + 3] = y1;
    const float cut = sigma_cutoff(opacities[i]);
    tighten_tile_rect(means2d[2 * i], means2d[2 * i + 1], conics[3 * i], conics[3 * i + 1], conics[3 * i + 2], cut,
                      (float)tile_size, x0, y0, x1, y1);
    pack_tile_rect(x0, y0, x1, y1, packed[2 * i], packed[2 * i + 1]);
  }
}
