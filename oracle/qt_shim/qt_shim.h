// Minimal stand-ins for the five Qt headers that /root/reference/src/sph.{h,cpp}
// include (sph.h:5-7, sph.cpp:8,16).  TEST INFRASTRUCTURE ONLY: lets the
// unmodified reference sources compile in place into oracle/_ref/ on a box
// without Qt5.  Nothing here is part of the product path.
#ifndef ORACLE_QT_SHIM_H
#define ORACLE_QT_SHIM_H

#include <cstdint>
#include <cstddef>
#include <chrono>
#include <vector>

#define Q_OBJECT
#define slots
#define signals public
#define emit

template <typename T>
class QList
{
public:
   void clear() { mItems.clear(); }
   void push_back(const T& v) { mItems.push_back(v); }
   bool isEmpty() const { return mItems.empty(); }
   int length() const { return (int)mItems.size(); }
   int count() const { return (int)mItems.size(); }
   int size() const { return (int)mItems.size(); }
   const T& operator[](int i) const { return mItems[(size_t)i]; }
   T& operator[](int i) { return mItems[(size_t)i]; }
   const T* data() const { return mItems.data(); }
private:
   std::vector<T> mItems;
};

class QMutex
{
public:
   void lock() {}
   void unlock() {}
};

class QThread
{
public:
   virtual ~QThread() {}
   virtual void run() {}
   void start() { run(); }   // synchronous: the harness has no GUI thread
   void quit() {}
   void wait() {}
};

class QElapsedTimer
{
public:
   void start() { mT0 = std::chrono::steady_clock::now(); }
   long long nsecsElapsed() const
   {
      return std::chrono::duration_cast<std::chrono::nanoseconds>(
         std::chrono::steady_clock::now() - mT0).count();
   }
private:
   std::chrono::steady_clock::time_point mT0;
};

#endif
