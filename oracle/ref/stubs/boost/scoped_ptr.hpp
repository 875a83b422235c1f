#pragma once
namespace boost { template <class T> class scoped_ptr { T *p; scoped_ptr(const scoped_ptr&); void operator=(const scoped_ptr&);
public: explicit scoped_ptr(T *q = 0) : p(q) {} ~scoped_ptr() { delete p; } T *get() const { return p; } T *operator->() const { return p; } T &operator*() const { return *p; } void reset(T *q = 0) { delete p; p = q; } }; }
