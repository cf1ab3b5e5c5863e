"""mosaicmagnifique_b200 -- B200-native photomosaic best-fit engine.

The product is libmosaic_b200.so (hand-written sm_100a CUDA behind the C ABI of include/mosaic_b200.h).
This package is the Python mirror of the reference's generator interface
(src/Photomosaic/PhotomosaicGeneratorBase.h:32-112) over that C ABI; it contains no compute of its own
and no CPU fallback: importing works anywhere, creating a generator needs the library and a CUDA device.
"""
from ._capi import MosaicError, capi, library_path  # noqa: F401
from .generator import (CIE76, CIEDE2000, RGB_EUCLIDEAN, CellGroup, CellShape, ColourDifference, ColourScheme,  # noqa: F401
                        PhotomosaicGenerator, load_mcs)
from .library import ImageLibrary  # noqa: F401

__all__ = ["PhotomosaicGenerator", "ImageLibrary", "CellShape", "CellGroup", "ColourDifference", "ColourScheme", "MosaicError", "capi",
           "library_path", "load_mcs", "RGB_EUCLIDEAN", "CIE76", "CIEDE2000"]
