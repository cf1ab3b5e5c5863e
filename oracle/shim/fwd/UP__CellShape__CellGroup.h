// forwards the reference's Windows-style include to its own header (found through -I$(REF)/src/CellShape)
#pragma once
#include "CellGroup.h"
