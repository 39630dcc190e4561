/* Build shim for the oracle/_ref recipe only.  The reference's base.h refuses
 * every platform except Windows and Apple; the recipe defines __APPLE__ and
 * this header answers the three macros base.h then asks for. */
#ifndef CAIRO_B200_REF_SHIM_TARGETCONDITIONALS_H
#define CAIRO_B200_REF_SHIM_TARGETCONDITIONALS_H
#define TARGET_OS_MAC 1
#define TARGET_OS_IPHONE 0
#define TARGET_IPHONE_SIMULATOR 0
#endif
