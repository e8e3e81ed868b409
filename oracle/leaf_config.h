/* TEST INFRASTRUCTURE ONLY: config.h stand-in when compiling against the reference headers in place */
#define X264_VERSION ""
#define X264_POINTVER "0.66.x"
