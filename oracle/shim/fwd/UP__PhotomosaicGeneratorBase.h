// forwards a back-end's Windows-style include (src/Photomosaic/<backend>/ -> "..\PhotomosaicGeneratorBase.h", as
// CUDA/CUDAPhotomosaicGenerator.h:24 does) to the reference's own header, found through -I$(REF)/src/Photomosaic
#pragma once
#include "PhotomosaicGeneratorBase.h"
