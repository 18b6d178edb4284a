// placeholder: path B lands next
#include "bundle_kernels.cuh"
#include "../../include/ptam_b200.h"
