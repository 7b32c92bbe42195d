// utils.hpp -- layout contract of the host API (same contract as reference include/utils.hpp:3-11):
// nine discrete velocities; population arrays are direction-innermost, then x, then y.
#pragma once

constexpr int Q = 9;

inline int INDEX(int x, int y, int i, int NX, int Q) { return i + Q * (x + NX * y); }
inline int INDEX(int x, int y, int NX) { return x + NX * y; }
