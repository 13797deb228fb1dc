// Minimal stand-in for <glog/logging.h>, which the reference includes from
// ITMLib/Engine/ITMTrackerFactory.h:7 but which is not installed in this image.
// TEST INFRASTRUCTURE ONLY: used when compiling the unmodified reference
// sources into oracle/_ref/ (see oracle/build_ref.py).
#pragma once
#include <iostream>
struct ItmB200NullLog {
  template <class T> ItmB200NullLog &operator<<(const T &) { return *this; }
  ItmB200NullLog &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
#define LOG(severity) ItmB200NullLog()
#define CHECK_NOTNULL(p) (p)
