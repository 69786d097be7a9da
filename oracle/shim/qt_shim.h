// TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-ins for the handful of Qt5
// types the reference's hot-path sources touch, so that the reference .cpp files
// can be compiled *verbatim from /root/reference* without Qt.
// Nothing here is product code; nothing here is copied from Qt.
#ifndef JAERO_ORACLE_QT_SHIM_H
#define JAERO_ORACLE_QT_SHIM_H
#include <vector>
#include <string>
#include <complex>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cctype>
#include <cassert>
#include <algorithm>
#include <initializer_list>
#include <iostream>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef int8_t qint8;   typedef uint8_t quint8;
typedef int16_t qint16; typedef uint16_t quint16;
typedef int32_t qint32; typedef uint32_t quint32;
typedef long long qint64; typedef unsigned long long quint64;
typedef double qreal;

#define Q_OBJECT
#define Q_UNUSED(x) (void)x;
#define signals public
#define slots
#define emit
#define SIGNAL(x) #x
#define SLOT(x) #x
#define Q_ENUM(x)
#define Q_DECL_OVERRIDE override

// Qt5 qglobal.h semantics: round half away from zero for >=0, and the odd
// negative-branch formula (qRound(-0.5)==0), see SURVEY.md P4.
inline int qRound(double d)
{ return d >= 0.0 ? int(d + 0.5) : int(d - double(int(d - 1)) + 0.5) + int(d - 1); }
template <class T> inline const T &qMin(const T &a, const T &b) { return (a < b) ? a : b; }
template <class T> inline const T &qMax(const T &a, const T &b) { return (a < b) ? b : a; }
template <class T> inline T qAbs(const T &t) { return t >= 0 ? t : -t; }

template <class T> class QVector
{
public:
    QVector() {}
    explicit QVector(int n) : v(n) {}
    QVector(int n, const T &t) : v(n, t) {}
    QVector(std::initializer_list<T> l) : v(l) {}
    int size() const { return (int)v.size(); }
    int count() const { return (int)v.size(); }
    int length() const { return (int)v.size(); }
    bool isEmpty() const { return v.empty(); }
    void resize(int n) { v.resize(n); }
    void reserve(int n) { v.reserve(n); }
    void clear() { v.clear(); }
    QVector<T> &fill(const T &t, int n = -1) { if (n >= 0) v.resize(n); std::fill(v.begin(), v.end(), t); return *this; }
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
    const T &at(int i) const { return v[i]; }
    void replace(int i, const T &t) { v[i] = t; }
    void push_back(const T &t) { v.push_back(t); }
    void append(const T &t) { v.push_back(t); }
    void append(const QVector<T> &o) { v.insert(v.end(), o.v.begin(), o.v.end()); }
    QVector<T> &operator<<(const T &t) { v.push_back(t); return *this; }
    QVector<T> &operator+=(const QVector<T> &o) { append(o); return *this; }
    T *data() { return v.data(); }
    const T *data() const { return v.data(); }
    const T *constData() const { return v.data(); }
    T &first() { return v.front(); }
    T &last() { return v.back(); }
    const T &first() const { return v.front(); }
    const T &last() const { return v.back(); }
    void removeFirst() { v.erase(v.begin()); }
    void removeLast() { v.pop_back(); }
    void remove(int i, int n = 1) { v.erase(v.begin() + i, v.begin() + i + n); }
    void removeAt(int i) { v.erase(v.begin() + i); }
    void insert(int i, const T &t) { v.insert(v.begin() + i, t); }
    QVector<T> mid(int pos, int len = -1) const
    {
        QVector<T> r;
        if (pos < 0) pos = 0;
        if (pos > size()) pos = size();
        int n = (len < 0 || pos + len > size()) ? size() - pos : len;
        r.v.assign(v.begin() + pos, v.begin() + pos + n);
        return r;
    }
    bool operator==(const QVector<T> &o) const { return v == o.v; }
    bool operator!=(const QVector<T> &o) const { return v != o.v; }
    typename std::vector<T>::iterator begin() { return v.begin(); }
    typename std::vector<T>::iterator end() { return v.end(); }
    typename std::vector<T>::const_iterator begin() const { return v.begin(); }
    typename std::vector<T>::const_iterator end() const { return v.end(); }
    std::vector<T> v;
};
template <class T> using QList = QVector<T>;

class QString;
class QByteArray
{
public:
    QByteArray() {}
    QByteArray(const char *s) : v(s, s + strlen(s)) {}
    QByteArray(const char *s, int n) : v(s, s + n) {}
    QByteArray(int n, char c) : v(n, c) {}
    int size() const { return (int)v.size(); }
    int length() const { return (int)v.size(); }
    int count() const { return (int)v.size(); }
    bool isEmpty() const { return v.empty(); }
    void resize(int n) { v.resize(n); }
    void reserve(int n) { v.reserve(n); }
    void clear() { v.clear(); }
    QByteArray &fill(char c, int n = -1) { if (n >= 0) v.resize(n); std::fill(v.begin(), v.end(), c); return *this; }
    char *data() { if (v.empty()) { v.reserve(1); } return v.data(); }
    const char *data() const { return v.data(); }
    const char *constData() const { return v.data(); }
    operator const char *() const { return v.data(); }
    char at(int i) const { return v[i]; }
    // Qt5: assigning through the non-const operator[] past the end grows the array (QByteRef); new bytes are zero here
    char &operator[](int i) { if (i >= (int)v.size()) v.resize(i + 1); return v[i]; }
    char operator[](int i) const { return v[i]; }
    void push_back(char c) { v.push_back(c); }
    QByteArray &append(char c) { v.push_back(c); return *this; }
    QByteArray &append(const QByteArray &o) { v.insert(v.end(), o.v.begin(), o.v.end()); return *this; }
    QByteArray &append(const char *s) { v.insert(v.end(), s, s + strlen(s)); return *this; }
    QByteArray &operator+=(char c) { return append(c); }
    QByteArray &operator+=(const QByteArray &o) { return append(o); }
    QByteArray &operator+=(const char *s) { return append(s); }
    QByteArray right(int n) const { if (n > size()) n = size(); return QByteArray(v.data() + size() - n, n); }
    QByteArray left(int n) const { if (n > size()) n = size(); return QByteArray(v.data(), n); }
    QByteArray mid(int pos, int len = -1) const
    {
        if (pos > size()) pos = size();
        int n = (len < 0 || pos + len > size()) ? size() - pos : len;
        return QByteArray(v.data() + pos, n);
    }
    void chop(int n) { if (n >= size()) v.clear(); else if (n > 0) v.resize(v.size() - n); }   // QByteArray::chop
    bool operator==(const QByteArray &o) const { return v == o.v; }
    std::vector<char> v;
};

class QChar
{
public:
    QChar(char ch = 0) : c(ch) {}
    char c;
};

class QString
{
public:
    QString() {}
    QString(const char *c) : s(c) {}
    QString(const std::string &c) : s(c) {}
    QString &operator+=(const QString &o) { s += o.s; return *this; }
    QString &operator+=(char c) { s += c; return *this; }
    QString operator+(const QString &o) const { return QString(s + o.s); }
    static QString number(double d) { return QString(std::to_string(d)); }
    static QString number(int d) { return QString(std::to_string(d)); }
    template <class A> QString arg(const A &) const { return *this; }
    // arg(value, fieldWidth, base, fillChar): replaces the %1 marker (the only one in the sources compiled here)
    QString arg(long long v, int width, int base, QChar fill) const
    {
        char digits[72]; int n = 0;
        unsigned long long u = (unsigned long long)(v < 0 ? -v : v);
        do { int d = (int)(u % (unsigned)base); digits[n++] = (char)(d < 10 ? '0' + d : 'a' + d - 10); u /= (unsigned)base; } while (u);
        std::string t; if (v < 0) t += '-';
        while (n) t += digits[--n];
        while ((int)t.size() < width) t.insert(t.begin(), fill.c);
        std::string r = s; size_t k = r.find("%1");
        if (k != std::string::npos) r.replace(k, 2, t);
        return QString(r);
    }
    QString &sprintf(const char *fmt, ...)
    {
        char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        s = buf; return *this;
    }
    QString toUpper() const { std::string r = s; for (char &c : r) c = (char)toupper((unsigned char)c); return QString(r); }
    int size() const { return (int)s.size(); }
    bool isEmpty() const { return s.empty(); }
    void clear() { s.clear(); }
    bool operator==(const QString &o) const { return s == o.s; }
    QByteArray toLatin1() const { return QByteArray(s.c_str()); }
    std::string s;
};
typedef QList<QString> QStringList;

struct QDebugSink
{
    template <class T> QDebugSink &operator<<(const T &) { return *this; }
};
inline QDebugSink qDebug() { return QDebugSink(); }

class QObject
{
public:
    explicit QObject(QObject *parent = 0) : parent_(parent) {}
    virtual ~QObject() {}
    QObject *parent() const { return parent_; }
    void setParent(QObject *p) { parent_ = p; }
    template <class... A> static bool connect(A...) { return true; }
    template <class... A> static bool disconnect(A...) { return true; }
    void deleteLater() {}
private:
    QObject *parent_;
};

class QIODevice : public QObject
{
public:
    enum OpenModeFlag { NotOpen = 0, ReadOnly = 1, WriteOnly = 2, ReadWrite = 3 };
    explicit QIODevice(QObject *parent = 0) : QObject(parent), open_(false) {}
    virtual bool open(int) { open_ = true; return true; }
    virtual void close() { open_ = false; }
    bool isOpen() const { return open_; }
    qint64 write(const QByteArray &b) { return writeData(b.data(), b.size()); }
    qint64 write(const char *d, qint64 n) { return writeData(d, n); }
    virtual qint64 readData(char *, qint64) { return 0; }
    virtual qint64 writeData(const char *, qint64 n) { return n; }
private:
    bool open_;
};

template <class T> class QPointer
{
public:
    QPointer() : p(0) {}
    QPointer(T *q) : p(q) {}
    QPointer<T> &operator=(T *q) { p = q; return *this; }
    bool isNull() const { return p == 0; }
    T *data() const { return p; }
    void clear() { p = 0; }
    T *operator->() const { return p; }
    operator T *() const { return p; }
private:
    T *p;
};

// Wall-clock GUI throttles are neutralised: elapsed() is always 0, so the
// "for looks" emits never fire and runs are deterministic.
class QElapsedTimer
{
public:
    void start() {}
    qint64 restart() { return 0; }
    qint64 elapsed() const { return 0; }
    bool isValid() const { return true; }
};
class QTimerEvent {};
class QTimer : public QObject
{
public:
    explicit QTimer(QObject *p = 0) : QObject(p) {}
    void start(int = 0) {}
    void stop() {}
};
class QFile {};
class QTextStream {};
class QDateTime {};
#endif
