// Build shim for the reference arm: this image has no libaio development package.
// Same ABI as the kernel's native AIO interface (linux/aio_abi.h); the five entry
// points the reference's RawAsync reader uses are thin syscall wrappers, exactly what
// libaio itself is.  Not part of the product; only used by baseline/build_reference.py.
#pragma once
#include <sys/syscall.h>
#include <unistd.h>
#include <errno.h>
#include <string.h>
#include <time.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct io_context* io_context_t;
struct iocb_common_shim {
  void* buf; unsigned long nbytes; long long offset; long long __pad3; unsigned flags; unsigned resfd;
};
struct iocb {                 // layout == struct iocb of linux/aio_abi.h (little endian, 64 bit)
  void* data;                      // aio_data
  unsigned key; unsigned aio_rw_flags;
  short aio_lio_opcode; short aio_reqprio; int aio_fildes;
  union { struct iocb_common_shim c; } u;
};
struct io_event { void* data; struct iocb* obj; long res; long res2; };
static inline int io_queue_init(int maxevents, io_context_t* ctxp) {
  *ctxp = 0;
  long r = syscall(SYS_io_setup, maxevents, ctxp);
  return r < 0 ? -errno : 0;
}
static inline int io_queue_release(io_context_t ctx) {
  long r = syscall(SYS_io_destroy, ctx); return r < 0 ? -errno : 0;
}
static inline int io_destroy(io_context_t ctx) { return io_queue_release(ctx); }
static inline void io_prep_pread(struct iocb* cb, int fd, void* buf, size_t count, long long offset) {
  memset(cb, 0, sizeof(*cb));
  cb->aio_fildes = fd; cb->aio_lio_opcode = 0 /* IOCB_CMD_PREAD */; cb->aio_reqprio = 0;
  cb->u.c.buf = buf; cb->u.c.nbytes = count; cb->u.c.offset = offset;
}
static inline int io_submit(io_context_t ctx, long nr, struct iocb* ios[]) {
  long r = syscall(SYS_io_submit, ctx, nr, ios); return r < 0 ? -errno : (int)r;
}
static inline int io_getevents(io_context_t ctx, long min_nr, long nr, struct io_event* events,
                               struct timespec* timeout) {
  long r = syscall(SYS_io_getevents, ctx, min_nr, nr, events, timeout); return r < 0 ? -errno : (int)r;
}
#ifdef __cplusplus
}
#endif
