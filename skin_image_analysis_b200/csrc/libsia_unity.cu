// The whole library is ONE translation unit: a single copy of the device-side watchdog word, no
// relocatable device code, one nvcc invocation (see skin_image_analysis_b200/build.py).
#include "core.cu"
#include "tmap.cu"
#include "counts.cu"
#include "preprocess.cu"
#include "preprocess_tc.cu"
#include "preprocess_tc2.cu"
#include "preprocess_mma.cu"
#include "preprocess_tv.cu"
#include "conv3x3.cu"
#include "conv1.cu"
#include "tail_cluster.cu"
#include "linear.cu"
