// oracle/stubs -- empty: SC.cpp uses nothing from OpenCV.
#pragma once
