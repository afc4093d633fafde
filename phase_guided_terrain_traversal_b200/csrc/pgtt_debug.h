// pgtt_debug.h - layout of the per-env record written by pgtt_debug_forward (floats).
#pragma once
#define DBG_XPOS 0        // [13][3]
#define DBG_XMAT 39       // [13][9]
#define DBG_XIPOS 156     // [13][3]
#define DBG_COM 195       // [3]
#define DBG_CINERT 198    // [13][10]
#define DBG_CDOF 328      // [18][6]
#define DBG_QM 436        // [18][18] dense
#define DBG_BIAS 760      // [18]
#define DBG_QS 778        // [18] qfrc_smooth
#define DBG_QAS 796       // [18] qacc_smooth
#define DBG_QACC 814      // [18]
#define DBG_CONTACT 832   // [8][16] dist pos3 frame9 mu leg box
#define DBG_EFC_D 960     // [44] limit rows 0..11 then contact rows
#define DBG_EFC_AREF 1004 // [44]
#define DBG_EFC_J 1048    // [44][18]
#define DBG_SENS 1840     // [49]
#define DBG_NITER 1889
#define DBG_ACTF 1890     // [12]
#define DBG_FOOT 1902     // [4][3]
#define DBG_END 1914
