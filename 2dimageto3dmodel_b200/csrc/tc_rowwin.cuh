// Row-window implicit-GEMM convolution (tc_conv3.cu), dispatched from b3d_conv2d_tf32 (tc_conv.cu) for stride-1 layers
// whose pixel tile is one 128-pixel row segment.
#pragma once
#include <cuda_runtime.h>

namespace b3d {
struct RowWinArgs {
    const float* x;      // [N, H, W, Cin]
    const float* wt;     // [kh*kw, Cout, Cin] tap-major (rows of a filter row are consecutive taps)
    const float* bias;
    float* out;          // pixel (n, y, x) -> out + ((n*OH + y + ooy)*OW + x + oox)*OC
    int N, H, W, Cin, Hout, Cout;
    int xlo, xhi;        // output columns of this launch: [xlo, xhi), (xhi - xlo) >= 128
    int kh, kw;
    int dy[5];           // input row of filter row r: y + dy[r]
    int dx0;             // first input column of the staged window: x + dx0
    int shift[5];        // window row shift of tap s of a filter row (dx[s] - dx0)
    int OH, OW, OC, ooy, oox;
    int osy, osx;        // output pixel (y, x) is written at (osy*y + ooy, osx*x + oox): 2 for the stride-2 dgrad parity classes
    int wtap0[5];        // weight tap (row of the tap-major array) of the first tap of filter row r
    int wtaps_total;     // taps in the weight array
    int wtap_step;       // tap-index distance between consecutive taps of a filter row (1, or 2 for the parity classes)
    float leaky;
    double* stats;       // nullable: [2][Cout] fp64 sums of the output (BN statistics), accumulated
};
// returns B3D_OK when launched, 1 when the geometry is not covered by a built variant (caller falls back)
int conv_rowwin_launch(const RowWinArgs& a, cudaStream_t st);
}  // namespace b3d
