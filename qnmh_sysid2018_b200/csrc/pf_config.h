// Launch configuration shared by the kernel instantiation units and the host side.
#pragma once
#ifndef PFB_BLOCK
#define PFB_BLOCK 512
#endif
#ifndef PFB_ITEMS
#define PFB_ITEMS 7
#endif
#ifndef PFB_MINB
#define PFB_MINB 1
#endif
#ifndef PFB_ITEMS_ALT
#define PFB_ITEMS_ALT 9
#endif
#ifndef PFB_SMOOTH_THREADS
#define PFB_SMOOTH_THREADS 128  // one warpgroup of smoother warps next to the PFB_BLOCK filter threads
#endif
#ifndef PFB_ITEMS_SMALL
#define PFB_ITEMS_SMALL 2  // small filters (<= 1024 particles per CTA): no masked work on empty items
#endif
