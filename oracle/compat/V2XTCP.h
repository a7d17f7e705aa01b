// oracle/compat/V2XTCP.h -- TEST INFRASTRUCTURE. Empty stand-in (Decision.cpp:4); the V2X TCP
// client is outside the reference and outside the hot path.
#pragma once
