/*
 * timing.h -- throughput / timer / log helpers with the reference's names
 * (/root/reference/lib/timing.h, lib/timing.cpp).
 */
#ifndef _TIMING_H_
#define _TIMING_H_

#include <fstream>
#include <iostream>
#include <string>

/** tab separated performance log; falls back to stderr without a file name (lib/timing.h:9-30) */
class Log {
public:
    Log(std::string filename) { if (!filename.empty()) fout.open(filename); }
    template <typename T>
    std::ostream& operator<<(T x)
    {
        std::ostream& o = fout.is_open() ? static_cast<std::ostream&>(fout) : std::cerr;
        o << x;
        return o;
    }
private:
    std::ofstream fout;
};

/** mebi-samples per second from a runtime in milliseconds (lib/timing.cpp:3-5) */
float throughput(float runtime, int pixels);

/** wall clock in milliseconds (lib/timing.cpp:23-28) */
unsigned long millisecond_timer(void);

#endif // _TIMING_H_
