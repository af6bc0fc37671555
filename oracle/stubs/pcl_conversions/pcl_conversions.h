// oracle/stubs -- in the ROS build this header pulls the C <math.h> into scope, which makes the
// unqualified atan()/sqrt() in SC.cpp:26-35,171 resolve to the FLOAT overloads (atanf, sqrtf).
#pragma once
#include <math.h>
