// forwards a back-end's "..\..\Other\Logger.h" (CUDA/CUDAPhotomosaicGenerator.cpp:28) to the stand-in logger
#pragma once
#include "..\Other\Logger.h"
