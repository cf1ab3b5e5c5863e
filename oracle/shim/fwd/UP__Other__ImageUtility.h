// forwards the reference's Windows-style include to its own header (path handed in by oracle/Makefile)
#pragma once
#include REF_IMAGEUTILITY_H
