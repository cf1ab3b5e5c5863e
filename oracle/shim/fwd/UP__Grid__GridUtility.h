// forwards the reference's Windows-style include to its own header (found through -I$(REF)/src/Grid)
#pragma once
#include "GridUtility.h"
