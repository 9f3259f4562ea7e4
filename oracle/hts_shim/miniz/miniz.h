/* TEST INFRASTRUCTURE -- abort-stubs for the four miniz calls of the reference's HTML-report zip
 * writer (output.cpp:5795-5846).  miniz is not in this image and that code is never reached by the
 * oracle build (oracle/ref_build.sh); the stubs only let output.cpp link. */
#ifndef BRQ_MINIZ_STUB_H
#define BRQ_MINIZ_STUB_H
#include <stdlib.h>
typedef struct mz_zip_archive { int unused; } mz_zip_archive;
typedef int mz_bool;
typedef unsigned int mz_uint;
typedef unsigned short mz_uint16;
typedef unsigned long long mz_uint64;
#define MZ_DEFAULT_COMPRESSION (-1)
static inline mz_bool mz_zip_writer_init_file(mz_zip_archive*, const char*, mz_uint64) { abort(); return 0; }
static inline mz_bool mz_zip_writer_add_file(mz_zip_archive*, const char*, const char*, const void*, mz_uint16, mz_uint) { abort(); return 0; }
static inline mz_bool mz_zip_writer_finalize_archive(mz_zip_archive*) { abort(); return 0; }
static inline mz_bool mz_zip_writer_end(mz_zip_archive*) { abort(); return 0; }
#endif
