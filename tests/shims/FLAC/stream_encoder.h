// TEST INFRASTRUCTURE: opaque FLAC encoder type, enough for `g++ -fsyntax-only` of the recorder module's headers
// (utils/wav.h:7,69); libFLAC is not installed in this image and nothing here is ever linked.
#pragma once
typedef struct FLAC__StreamEncoder FLAC__StreamEncoder;
