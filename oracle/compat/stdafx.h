// oracle/compat/stdafx.h -- TEST INFRASTRUCTURE. Stand-in for the MFC precompiled header the
// reference includes first (Planning.cpp:1, Decision.cpp:1).  Supplies the Win32 names the two
// translation units use, so that the UNMODIFIED reference sources compile with g++.
#pragma once
// every std header first: min/max become macros at the bottom (MSVC <windows.h> behaviour the
// reference relies on: max(double,int) Decision.cpp:375, min(WORD,int) Decision.cpp:581)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <string>
#include <vector>

typedef unsigned char BYTE;
typedef unsigned short WORD;
typedef unsigned int DWORD;
typedef unsigned int UINT;
typedef int INT;
typedef int BOOL;
typedef float FLOAT;
typedef double DOUBLE;
typedef void* LPVOID;
typedef void* HANDLE;
typedef DWORD (*LPTHREAD_START_ROUTINE)(LPVOID);
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
union LARGE_INTEGER { long long QuadPart; };

// implemented by oracle/ref_harness.cpp (cooperative scheduler, deterministic clock)
extern "C" {
DWORD WaitForSingleObject(HANDLE h, DWORD timeout_ms);
BOOL SetEvent(HANDLE h);
BOOL QueryPerformanceCounter(LARGE_INTEGER* t);
HANDLE CreateThread(void*, size_t, LPTHREAD_START_ROUTINE, LPVOID, DWORD, DWORD*);
void Sleep(DWORD ms);
}
int AfxMessageBox(const char* msg);
void* AfxGetApp();

class CCriticalSection {
public:
    void Lock() {}
    void Unlock() {}
};

using namespace std;
#undef min
#undef max
#define min(a, b) (((a) < (b)) ? (a) : (b))
#define max(a, b) (((a) > (b)) ? (a) : (b))
