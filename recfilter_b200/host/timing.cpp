/* timing.cpp -- see include/timing.h (restates /root/reference/lib/timing.cpp) */
#include <timing.h>
#include <chrono>

float throughput(float runtime_ms, int pixels)
{
    return (float(pixels) * 1000.0f) / (runtime_ms * 1024.0f * 1024.0f);
}

unsigned long millisecond_timer(void)
{
    using namespace std::chrono;
    return (unsigned long)duration_cast<milliseconds>(system_clock::now().time_since_epoch()).count();
}
