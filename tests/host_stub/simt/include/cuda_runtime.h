// TEST INFRASTRUCTURE ONLY: stands in for <cuda_runtime.h> when the kernels are compiled for the host.
#pragma once
#include "../simt_host.h"
