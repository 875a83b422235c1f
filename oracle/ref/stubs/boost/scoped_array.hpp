#pragma once
namespace boost { template <class T> class scoped_array { T *p; scoped_array(const scoped_array&); void operator=(const scoped_array&);
public: explicit scoped_array(T *q = 0) : p(q) {} ~scoped_array() { delete[] p; } T *get() const { return p; } T &operator[](long i) const { return p[i]; } void reset(T *q = 0) { delete[] p; p = q; } }; }
