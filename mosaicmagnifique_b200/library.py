"""ImageLibrary (src/ImageLibrary/ImageLibrary.{h,cpp}) over the C ABI: the container either side of the generator's
`setLibrary`. Crop + resize of every added image run on the GPU (`mosaic_library_ingest`); the container bookkeeping
(names, random insertion index, .mil load / save) is host logic and lives here."""
import random

import numpy as np

from ._capi import MosaicError, capi
from .formats import load_mil, save_mil


class ImageLibrary:
    """Same method names and behaviour as the reference class (ImageLibrary.h:9-66)."""

    def __init__(self, imageSize: int, device: int = 0, seed=None):
        self._size = int(imageSize)
        self._device = device
        self._names = []
        self._originals = []  # square-cropped, resized to the size that was current when they were added
        self._resized = []
        # the reference draws the insertion index from std::random_device (ImageLibrary.cpp:75-78); a seed makes it repeatable
        self._rng = random.Random(seed)

    def __eq__(self, other):  # operator==, ImageLibrary.cpp:15-39: image size, image count, resized images (not the names)
        return (isinstance(other, ImageLibrary) and self._size == other._size and len(self._resized) == len(other._resized)
                and all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(self._resized, other._resized)))

    def _ingest(self, im: np.ndarray, size: int) -> np.ndarray:
        im = np.asarray(im)
        if im.size == 0:
            raise ValueError("t_im was empty.")  # std::invalid_argument, ImageLibrary.cpp:65-66
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise ValueError("library images are 8-bit BGR")
        if im.strides[2] != 1 or im.strides[1] != 3:
            im = np.ascontiguousarray(im)
        out = np.empty((size, size, 3), np.uint8)
        rc = capi().mosaic_library_ingest(self._device, im.ctypes.data, im.shape[0], im.shape[1], im.strides[0], size,
                                          out.ctypes.data)
        if rc:
            raise MosaicError(rc, "library ingest failed")
        return out

    def setImageSize(self, size: int):
        """ImageLibrary.cpp:42-52: every stored original is resized again (EXACT) to the new size."""
        size = int(size)
        if size == self._size:
            return
        self._size = size
        self._resized = [self._ingest(o, size) for o in self._originals]

    def getImageSize(self) -> int:
        return self._size

    def addImage(self, im: np.ndarray, name: str = "") -> int:
        """ImageLibrary.cpp:62-86: centre crop to square, resize to the library size, insert at a random index."""
        img = self._ingest(im, self._size)
        index = self._rng.randint(0, len(self._originals))
        self._names.insert(index, name)
        self._originals.insert(index, img)  # addImageInternal stores the SAME resized image twice (:240-244)
        self._resized.insert(index, img)
        return index

    def getNames(self):
        return self._names

    def getImages(self):
        return self._resized

    def asArray(self) -> np.ndarray:
        """[N][size][size][3] contiguous, the form PhotomosaicGenerator.setLibrary takes."""
        if not self._resized:
            return np.zeros((0, self._size, self._size, 3), np.uint8)
        return np.ascontiguousarray(np.stack(self._resized))

    def removeAtIndex(self, index: int):
        del self._names[index], self._originals[index], self._resized[index]

    def clear(self):
        self._names, self._originals, self._resized = [], [], []

    def saveToFile(self, filename: str):
        """ImageLibrary.cpp:117-154 (.mil version 6, PNG-encoded images)."""
        if not filename:
            raise ValueError("No filename")
        save_mil(filename, self.asArray(), self._names)

    def loadFromFile(self, filename: str):
        """ImageLibrary.cpp:157-236: appends the file's images; the file's image size becomes the library's."""
        if not filename:
            raise ValueError("No filename")
        images, names, size = load_mil(filename)
        self._size = size
        for img, name in zip(images, names):
            self._names.append(name)
            self._originals.append(img)
            self._resized.append(img)
