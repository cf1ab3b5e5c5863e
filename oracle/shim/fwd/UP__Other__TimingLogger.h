// Stand-in for src/Other/TimingLogger.h: the timing scopes are no-ops.
#pragma once
#include <QString>
class TimingLogger {
public:
    void StartTiming(const QString &) {}
    void StopTiming(const QString &) {}
    void StopAllTiming() {}
    void LogTiming() {}
};
