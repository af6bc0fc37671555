// oracle/stubs -- empty: SC.cpp never uses VoxelGrid.
#pragma once
