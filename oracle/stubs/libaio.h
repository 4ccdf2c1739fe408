/* TEST INFRASTRUCTURE (oracle/stubs) — declaration-only stand-in for <libaio.h>, which is not installed here.
 * The reference's aligned_file_reader.h only needs the io_context_t type name; the in-memory reader of
 * diskann_flash_harness.cpp never submits asynchronous I/O. */
#ifndef LB2_STUB_LIBAIO_H
#define LB2_STUB_LIBAIO_H
typedef struct lb2_stub_io_context* io_context_t;
#endif
