// oracle/compat/GAC_Auotpilot_DPDlg.h -- TEST INFRASTRUCTURE. Empty stand-in (Decision.cpp:6).
#pragma once
