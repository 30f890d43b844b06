// Row-marching 3x3 convolution, build with cp.async row producers (two warps gather the 130-pixel window of the line 16
// bytes at a time: any mix of images and pad pixels) -- for images narrower than 129 pixels.  See conv_row_impl.cuh.
#define LV_ROW_CPASYNC 1
#define LV_ROW_NS row_cp
#define LV_ROW_ENTRY conv3x3_row_chain_cp
#include "conv_row_impl.cuh"
