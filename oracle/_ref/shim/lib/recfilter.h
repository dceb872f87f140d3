#include <recfilter.h>
