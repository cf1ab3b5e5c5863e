// B200PhotomosaicGenerator.h -- the reference-side binding: a third back-end next to CPUPhotomosaicGenerator /
// CUDAPhotomosaicGenerator (src/Photomosaic/CPUPhotomosaicGenerator.h:25-33, src/Photomosaic/CUDA/CUDAPhotomosaicGenerator.h:27-60).
// A maintainer drops this file and the .cpp into src/Photomosaic/B200/ and links libmosaic_b200.so; like the existing
// back-ends it overrides generateBestFits() only -- setters, getBestFits(), buildPhotomosaic(), cancel(), progress(int) stay
// the reference's own (PhotomosaicGeneratorBase.h:32-112).
// In this repo the pair is compiled against the reference's real PhotomosaicGeneratorBase.h (Qt / OpenCV replaced by the
// stand-ins of oracle/shim) and driven through the base-class API by tests/test_gpu_dropin.py.
#pragma once
#include "..\PhotomosaicGeneratorBase.h"

#include <mosaic_b200.h>  // include/mosaic_b200.h of this repo

class B200PhotomosaicGenerator : public PhotomosaicGeneratorBase
{
    Q_OBJECT
public:
    explicit B200PhotomosaicGenerator(const int device = 0);
    ~B200PhotomosaicGenerator() override;

    //Generate best fits for Photomosaic cells
    //Returns true if successful
    bool generateBestFits() override;

    //Text of the last engine error (the reference shows a modal box, CUDAUtility.h:34-62)
    const char *lastError() const;

private:
    mosaic_generator *m_engine = nullptr;
};
