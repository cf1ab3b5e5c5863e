// Stand-in for src/Other/Logger.h + Utility.h: logging is a no-op, MessageBox::critical records that it was called.
#pragma once
#include "qt_standins.h"
#define LogInfo(x) ((void)0)
#define LogCritical(x) ((void)0)
extern int g_ref_message_boxes;
struct MessageBox {
    static void critical(void *, const QString &, const QString &) { ++g_ref_message_boxes; }
};
