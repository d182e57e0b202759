// TEST INFRASTRUCTURE: opaque lame handle for `g++ -fsyntax-only` of utils/wav.h:8,65 (libmp3lame is not installed here).
#pragma once
typedef struct lame_global_struct* lame_t;
