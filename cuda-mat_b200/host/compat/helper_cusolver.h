// helper_cusolver.h shim: example.cpp:351-353 only needs second() (helper_cusolver.h:148-153).
#pragma once
#include <time.h>
inline double second(void) {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
