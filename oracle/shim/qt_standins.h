// Stand-ins for the Qt pieces the reference's generator / cell-shape / image-library sources mention. Not Qt code.
// Test infrastructure: lets those sources compile UNMODIFIED into oracle/_ref (see oracle/Makefile).
//   QObject + the moc keywords, tr()            -- PhotomosaicGeneratorBase.h
//   QString, QByteArray, QDataStream, QFile     -- CellShape.cpp / ImageLibrary.cpp (.mcs / .mil containers): the wire format is
//                                                  Qt's documented QDataStream encoding (big-endian integers; QString = u32
//                                                  byte length + UTF-16BE, 0xFFFFFFFF for a null string; QByteArray = u32 length
//                                                  + bytes; bool = one byte), which is what makes the reference's own
//                                                  loadFromFile / saveToFile usable on the real Cells/*.mcs files here.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

typedef uint32_t quint32;
typedef int32_t qint32;
typedef uint16_t quint16;

class QString {
    std::string s_;  // UTF-8
    bool null_ = true;
public:
    QString() {}
    QString(const char *s) : s_(s ? s : ""), null_(s == nullptr) {}
    static QString fromStdString(const std::string &s)
    {
        QString q;
        q.s_ = s;
        q.null_ = false;
        return q;
    }
    bool isNull() const { return null_; }
    bool isEmpty() const { return s_.empty(); }
    std::string toStdString() const { return s_; }
    int compare(const QString &o) const { return s_.compare(o.s_); }
    bool operator==(const QString &o) const { return s_ == o.s_; }
    bool operator!=(const QString &o) const { return s_ != o.s_; }
    // UTF-16 code units of the string (surrogate pairs above the BMP)
    std::vector<quint16> utf16() const
    {
        std::vector<quint16> out;
        for (size_t i = 0; i < s_.size();) {
            const unsigned char c = (unsigned char)s_[i];
            uint32_t cp;
            int n;
            if (c < 0x80) { cp = c; n = 1; }
            else if ((c >> 5) == 6) { cp = c & 31; n = 2; }
            else if ((c >> 4) == 14) { cp = c & 15; n = 3; }
            else { cp = c & 7; n = 4; }
            for (int k = 1; k < n && i + k < s_.size(); ++k)
                cp = (cp << 6) | ((unsigned char)s_[i + k] & 63);
            i += n;
            if (cp >= 0x10000) {
                cp -= 0x10000;
                out.push_back((quint16)(0xD800 + (cp >> 10)));
                out.push_back((quint16)(0xDC00 + (cp & 0x3FF)));
            } else
                out.push_back((quint16)cp);
        }
        return out;
    }
    static QString fromUtf16(const std::vector<quint16> &u)
    {
        std::string s;
        for (size_t i = 0; i < u.size(); ++i) {
            uint32_t cp = u[i];
            if (cp >= 0xD800 && cp < 0xDC00 && i + 1 < u.size())
                cp = 0x10000 + ((cp - 0xD800) << 10) + (u[++i] - 0xDC00);
            if (cp < 0x80) s += (char)cp;
            else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
            else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
            else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 63)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
        }
        return fromStdString(s);
    }
};
#ifndef Q_FUNC_INFO
#define Q_FUNC_INFO ""
#endif

class QByteArray {
    std::vector<char> d_;
public:
    QByteArray() {}
    QByteArray(const char *p, int n) : d_(p, p + (n > 0 ? n : 0)) {}
    static QByteArray fromRawData(const char *p, int n) { return QByteArray(p, n); }
    char *data() { return d_.data(); }
    const char *data() const { return d_.data(); }
    int size() const { return (int)d_.size(); }
    std::vector<char>::const_iterator cbegin() const { return d_.cbegin(); }
    std::vector<char>::const_iterator cend() const { return d_.cend(); }
    void resize(int n) { d_.resize((size_t)n); }
};

class QIODevice {
public:
    enum OpenModeFlag { NotOpen = 0, ReadOnly = 1, WriteOnly = 2, ReadWrite = 3 };
    virtual ~QIODevice() {}
    virtual bool readBytes(void *dst, size_t n) = 0;
    virtual bool writeBytes(const void *src, size_t n) = 0;
};

class QFile : public QIODevice {
    std::string name_;
    FILE *f_ = nullptr;
    int mode_ = 0;
public:
    explicit QFile(const QString &name) : name_(name.toStdString()) {}
    ~QFile() override { close(); }
    bool open(int mode)
    {
        close();
        f_ = std::fopen(name_.c_str(), mode == ReadOnly ? "rb" : "wb");
        mode_ = f_ ? mode : 0;
        return f_ != nullptr;
    }
    bool isReadable() const { return f_ && (mode_ & ReadOnly); }
    bool isWritable() const { return f_ && (mode_ & WriteOnly); }
    void close()
    {
        if (f_)
            std::fclose(f_);
        f_ = nullptr;
        mode_ = 0;
    }
    bool readBytes(void *dst, size_t n) override { return f_ && std::fread(dst, 1, n, f_) == n; }
    bool writeBytes(const void *src, size_t n) override { return f_ && std::fwrite(src, 1, n, f_) == n; }
};

class QDataStream {
    QIODevice *dev_;
    bool ok_ = true;
    template <typename T> void put_be(T v)
    {
        unsigned char b[sizeof(T)];
        for (size_t i = 0; i < sizeof(T); ++i)
            b[i] = (unsigned char)((uint64_t)v >> (8 * (sizeof(T) - 1 - i)));
        ok_ = dev_->writeBytes(b, sizeof(T)) && ok_;
    }
    template <typename T> T get_be()
    {
        unsigned char b[sizeof(T)] = {};
        ok_ = dev_->readBytes(b, sizeof(T)) && ok_;
        uint64_t v = 0;
        for (size_t i = 0; i < sizeof(T); ++i)
            v = (v << 8) | b[i];
        return (T)v;
    }
public:
    enum Version { Qt_5_0 = 13 };
    explicit QDataStream(QIODevice *d) : dev_(d) {}
    // raw access for the free operators below (Qt declares the QString / QByteArray operators outside the class too, which is
    // why `customStream >> qstring` finds them although CustomQDataStream has member operators of its own)
    void writeRaw(const void *p, size_t n) { ok_ = dev_->writeBytes(p, n) && ok_; }
    void readRaw(void *p, size_t n) { ok_ = dev_->readBytes(p, n) && ok_; }
    virtual ~QDataStream() {}
    void setVersion(int) {}
    bool ok() const { return ok_; }
    QDataStream &operator<<(quint32 v) { put_be<quint32>(v); return *this; }
    QDataStream &operator<<(qint32 v) { put_be<quint32>((quint32)v); return *this; }
    QDataStream &operator<<(bool v) { put_be<unsigned char>(v ? 1 : 0); return *this; }
    QDataStream &operator>>(quint32 &v) { v = get_be<quint32>(); return *this; }
    QDataStream &operator>>(qint32 &v) { v = (qint32)get_be<quint32>(); return *this; }
    QDataStream &operator>>(bool &v) { v = get_be<unsigned char>() != 0; return *this; }
};

inline QDataStream &operator<<(QDataStream &st, const QString &s)
{
    if (s.isNull())
        return st << (quint32)0xFFFFFFFFu;
    const std::vector<quint16> u = s.utf16();
    st << (quint32)(u.size() * 2);
    for (quint16 c : u) {
        const unsigned char b[2] = {(unsigned char)(c >> 8), (unsigned char)(c & 255)};
        st.writeRaw(b, 2);
    }
    return st;
}
inline QDataStream &operator>>(QDataStream &st, QString &s)
{
    quint32 n = 0;
    st >> n;
    if (n == 0xFFFFFFFFu || !st.ok()) {
        s = QString();
        return st;
    }
    std::vector<quint16> u(n / 2);
    for (auto &c : u) {
        unsigned char b[2] = {0, 0};
        st.readRaw(b, 2);
        c = (quint16)((b[0] << 8) | b[1]);
    }
    s = QString::fromUtf16(u);
    return st;
}
inline QDataStream &operator<<(QDataStream &st, const QByteArray &a)
{
    st << (quint32)a.size();
    if (a.size())
        st.writeRaw(a.data(), (size_t)a.size());
    return st;
}
inline QDataStream &operator>>(QDataStream &st, QByteArray &a)
{
    quint32 n = 0;
    st >> n;
    a = QByteArray();
    if (n == 0xFFFFFFFFu || !st.ok())
        return st;
    a.resize((int)n);
    if (n)
        st.readRaw(a.data(), n);
    return st;
}

// qDebug(): a sink
struct QDebugSink {
    template <typename T> QDebugSink &operator<<(const T &) { return *this; }
};
inline QDebugSink qDebug() { return QDebugSink(); }

// GUI types ImageUtility.h mentions (progress bar of batchResizeMat, pixmap conversion for previews): inert
class QProgressBar {
    int v_ = 0;
public:
    void setMaximum(int) {}
    void setValue(int v) { v_ = v; }
    int value() const { return v_; }
    void setVisible(bool) {}
};
class QImage {
public:
    enum Format { Format_RGB888 = 13 };
    QImage() {}
    QImage(const unsigned char *, int, int, int, Format) {}
    QImage rgbSwapped() const { return *this; }
};
class QPixmap {
public:
    static QPixmap fromImage(const QImage &) { return QPixmap(); }
};

class QObject {
public:
    virtual ~QObject() {}
    static QString tr(const char *s) { return QString(s); }
};
#define Q_OBJECT
#define slots
#define signals public
#define emit
