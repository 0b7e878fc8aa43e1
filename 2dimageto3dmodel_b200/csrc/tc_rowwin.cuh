// Row-window implicit-GEMM convolution (tc_conv3.cu), dispatched from b3d_conv2d_tf32 (tc_conv.cu) for stride-1 layers
// whose pixel tile is one 128-pixel row segment.
#pragma once
#include <cuda_runtime.h>

namespace b3d {
struct RowWinArgs {
    const float* x;      // [N, H, W, Cin]
    const float* wt;     // [kh*kw, Cout, Cin] tap-major (rows of a filter row are consecutive taps)
    const float* bias;
    float* out;          // pixel (n, y, x) of class c -> out + ((n*OH + osy*y + ooy[c])*OW + osx*x + oox[c])*OC
    int N, H, W, Cin, Hout, Cout;
    int xlo, xhi;        // output columns of this launch: [xlo, xhi), (xhi - xlo) >= 128
    int kh, kw;
    int ncls;            // >= 1 output classes in one launch (stride-2 input-gradient parity classes): per-class rows / taps / offsets
    int dy[4][5];        // class c: input row of filter row r: y + dy[c][r]
    int dx0;             // first input column of the staged window: x + dx0
    int shift[5];        // window row shift of tap s of a filter row (dx[s] - dx0)
    int OH, OW, OC, ooy[4], oox[4];
    int osy, osx;        // output pixel (y, x) of class c is written at (osy*y + ooy[c], osx*x + oox[c]): 2 for the parity classes
    int wtap0[4][5];     // class c: weight tap (row of the tap-major array) of the first tap of filter row r
    int wtaps_total;     // taps in the weight array
    int wtap_step;       // tap-index distance between consecutive taps of a filter row (1, or 2 for the parity classes)
    float leaky;
    double* stats;       // nullable: [2][Cout] fp64 sums of the output (BN statistics), accumulated
    const float* mask;   // nullable: LeakyReLU adjoint fused into the epilogue (b3d_conv_opts in include/b3d.h)
    float mslope;
    int stats_sum;       // statistics: sums only
    int xpitch;          // pixels per image row of x in memory (0 = W)
};
// returns B3D_OK when launched, 1 when the geometry is not covered by a built variant (caller falls back)
int conv_rowwin_launch(const RowWinArgs& a, cudaStream_t st);
}  // namespace b3d
