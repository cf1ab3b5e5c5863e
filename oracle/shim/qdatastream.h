// stand-in header: everything Qt-like lives in qt_standins.h (same directory)
#pragma once
#include "qt_standins.h"
