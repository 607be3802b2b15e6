/*
 * Panama FFM (java.lang.foreign, JDK 22+) binding of libjtb200.so -- the C ABI declared in include/jtb200.h.
 * SOURCE ONLY: no JDK exists in the build image or on the GPU box, so this file has never been compiled.
 * It is the binding a JTransforms maintainer adds so that the org.jtransforms classes keep their public
 * signatures while ConcurrencyUtils.submit/waitForCompletion dispatch is replaced by GPU launches.
 */
package org.jtransforms.b200;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import static java.lang.foreign.ValueLayout.*;

public final class Jtb200 {
    public static final int FFT = 0, DCT = 1, DST = 2, DHT = 3;
    public static final int F64 = 0, F32 = 1;
    public static final int C2C_FORWARD = 0, C2C_INVERSE = 1, R2C_PACKED = 2, R2C_FULL = 3, C2R_PACKED = 4,
                            C2R_FULL = 5, R2R_FORWARD = 6, R2R_INVERSE = 7;
    private static final int ERR_ARG = 1;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB =
        SymbolLookup.libraryLookup(System.getProperty("jtb200.library", "libjtb200.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd, Linker.Option... opts) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), fd, opts);
    }

    // int jtb_plan_create(jtb_plan** out, int kind, int prec, int rank, const int64_t* dims, int device)
    private static final MethodHandle PLAN_CREATE = h("jtb_plan_create",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle PLAN_DESTROY = h("jtb_plan_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    // int jtb_exec(jtb_plan*, int op, void* host_a, int64_t offa, int scale)
    // Linker.Option.critical(true): the heap array is passed without a copy (pinned for the call), which is what
    // makes the double[] overloads in-place like the reference (fft/DoubleFFT_1D.java:243-263).
    private static final MethodHandle EXEC = h("jtb_exec",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT), Linker.Option.critical(true));
    private static final MethodHandle EXEC_BATCH = h("jtb_exec_batch",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG, JAVA_INT),
        Linker.Option.critical(true));
    private static final MethodHandle LAST_ERROR = h("jtb_last_error", FunctionDescriptor.of(ADDRESS));
    // int jtb_host_register(void* p, int64_t bytes) / int jtb_host_unregister(void* p): page-lock off-heap storage
    // (the memory behind a DoubleLargeArray / FloatLargeArray) once so that every later exec copies at DMA speed
    private static final MethodHandle HOST_REGISTER = h("jtb_host_register", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));
    private static final MethodHandle HOST_UNREGISTER = h("jtb_host_unregister", FunctionDescriptor.of(JAVA_INT, ADDRESS));

    /** Page-locks a native segment until the returned handle is closed. */
    public static AutoCloseable pin(MemorySegment seg) {
        try {
            check((int) HOST_REGISTER.invokeExact(seg, seg.byteSize()));
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        return () -> {
            try {
                check((int) HOST_UNREGISTER.invokeExact(seg));
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        };
    }

    /** One jtb_plan*: immutable, thread-safe, freed by close() (or a Cleaner registered by the owner). */
    public static final class Plan implements AutoCloseable {
        private MemorySegment handle;

        public Plan(int kind, int prec, long... dims) {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment out = a.allocate(ADDRESS);
                MemorySegment d = a.allocateFrom(JAVA_LONG, dims);
                check((int) PLAN_CREATE.invokeExact(out, kind, prec, dims.length, d, 0));
                handle = out.get(ADDRESS, 0);
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }

        /** in place on a Java heap array: a[offa ...] */
        public void exec(int op, double[] a, long offa, boolean scale) {
            try {
                check((int) EXEC.invokeExact(handle, op, MemorySegment.ofArray(a), offa, scale ? 1 : 0));
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }

        public void exec(int op, float[] a, long offa, boolean scale) {
            try {
                check((int) EXEC.invokeExact(handle, op, MemorySegment.ofArray(a), offa, scale ? 1 : 0));
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }

        /** off-heap storage (DoubleLargeArray.isLarge(): pass its native address; > 2^31 elements are fine) */
        public void exec(int op, MemorySegment a, long offa, boolean scale) {
            try {
                check((int) EXEC.invokeExact(handle, op, a, offa, scale ? 1 : 0));
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }

        @Override public void close() {
            try {
                if (handle != null) { int rc = (int) PLAN_DESTROY.invokeExact(handle); handle = null; }
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }
    }

    /** Argument errors keep the reference's exception type and text; everything else is IllegalStateException. */
    static void check(int status) {
        if (status == 0) return;
        String msg;
        try {
            msg = ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(512).getString(0);
        } catch (Throwable t) {
            msg = "libjtb200 error " + status;
        }
        if (status == ERR_ARG) throw new IllegalArgumentException(msg);
        throw new IllegalStateException("libjtb200 error " + status + ": " + msg);
    }

    private Jtb200() { }
}
