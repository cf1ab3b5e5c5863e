// Stand-ins for the Qt pieces the reference's generator headers mention (QObject, the moc keywords, tr). Not Qt code.
// Test infrastructure: lets CPUPhotomosaicGenerator.cpp compile UNMODIFIED into oracle/_ref (see oracle/Makefile).
#pragma once
#include <QString>
class QObject {
public:
    virtual ~QObject() {}
    static QString tr(const char *s) { return QString(s); }
};
#define Q_OBJECT
#define slots
#define signals public
#define emit
