// padded ranks with compiled kernels: even values to 32, then multiples of 8 to 64
#pragma once
#define VB_RP_LIST(F) \
    F(2) F(4) F(6) F(8) F(10) F(12) F(14) F(16) F(18) F(20) F(22) F(24) F(26) F(28) F(30) F(32) \
    F(40) F(48) F(56) F(64)
