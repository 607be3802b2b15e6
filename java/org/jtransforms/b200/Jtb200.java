/*
 * Panama FFM (java.lang.foreign, JDK 22+) binding of libjtb200.so -- the C ABI declared in include/jtb200.h.
 * SOURCE ONLY: no JDK exists in the build image or on the GPU box, so this file has never been compiled.
 * It is the binding a JTransforms maintainer adds so that the org.jtransforms classes keep their public
 * signatures while ConcurrencyUtils.submit/waitForCompletion dispatch is replaced by GPU launches.
 */
package org.jtransforms.b200;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import org.visnow.jlargearrays.DoubleLargeArray;
import org.visnow.jlargearrays.FloatLargeArray;
import org.visnow.jlargearrays.LargeArrayUtils;
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
    // int jtb_exec_n(jtb_plan*, int op, void* host_a, int64_t a_length, int64_t offa, int scale): the array length travels
    // with the call; "array too short" comes back as JTB_ERR_ARG and is rethrown as ArrayIndexOutOfBoundsException
    // (what the reference's unchecked loops raise, fft/DoubleFFT_1D.java:243-263) instead of touching the heap past the end
    private static final MethodHandle EXEC_N = h("jtb_exec_n",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT), Linker.Option.critical(true));
    private static final MethodHandle PLAN_ELEMENTS = h("jtb_plan_elements", FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_INT));
    private static final MethodHandle PLAN_SET_DEVICES = h("jtb_plan_set_devices",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
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

        /** number of array elements the op reads and writes for one transform (jtb_plan_elements) */
        public long elements(int op) {
            try {
                return (long) PLAN_ELEMENTS.invokeExact(handle, op);
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }

        /** multi-GPU plan: host-array 3-D transforms are slab-decomposed over `devices`, batches are split (jtb_plan_set_devices) */
        public void setDevices(int... devices) {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment d = a.allocateFrom(JAVA_INT, devices);
                check((int) PLAN_SET_DEVICES.invokeExact(handle, devices.length, d));
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }

        private void execN(int op, MemorySegment a, long length, long offa, boolean scale) {
            if (offa < 0 || offa + elements(op) > length)
                throw new ArrayIndexOutOfBoundsException("array of " + length + " elements is too short: the transform touches ["
                                                         + offa + ", " + (offa + elements(op)) + ")");
            try {
                check((int) EXEC_N.invokeExact(handle, op, a, length, offa, scale ? 1 : 0));
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }

        /** in place on a Java heap array: a[offa ...]; the heap array is pinned for the call, not copied */
        public void exec(int op, double[] a, long offa, boolean scale) { execN(op, MemorySegment.ofArray(a), a.length, offa, scale); }

        public void exec(int op, float[] a, long offa, boolean scale) { execN(op, MemorySegment.ofArray(a), a.length, offa, scale); }

        /** off-heap storage of `length` elements (no 2^31 limit) */
        public void exec(int op, MemorySegment a, long length, long offa, boolean scale) { execN(op, a, length, offa, scale); }

        @Override public void close() {
            try {
                if (handle != null) { int rc = (int) PLAN_DESTROY.invokeExact(handle); handle = null; }
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        }
    }

    // ------------------------------------------------------------------------------------------------ LargeArray bridge
    // DoubleLargeArray / FloatLargeArray (fft/DoubleFFT_1D.java:280-304): arrays below the JLargeArrays threshold are
    // plain Java arrays (getData()) and take the pinned heap path; larger ones are staged through one native block in
    // CHUNK-element pieces with LargeArrayUtils.arraycopy -- the only bulk accessor the reference itself uses -- so
    // nothing here depends on JLargeArrays internals.  (A maintainer with access to the off-heap address can hand
    // Plan.exec(op, MemorySegment, ...) that address instead and skip the staging copy.)
    private static final int CHUNK = 1 << 24;

    public static void execLarge(Plan plan, int op, DoubleLargeArray a, long offa, boolean scale) {
        if (a.isConstant()) throw new IllegalArgumentException("The data array is constant.");
        if (!a.isLarge() && offa < Integer.MAX_VALUE) { plan.exec(op, a.getData(), offa, scale); return; }
        final long n = plan.elements(op);
        if (offa < 0 || offa + n > a.length()) throw new ArrayIndexOutOfBoundsException("array too short");
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment seg = arena.allocate(8L * n, 64);
            DoubleLargeArray tmp = new DoubleLargeArray((long) CHUNK, false);
            for (long p = 0; p < n; p += CHUNK) {
                final long len = Math.min((long) CHUNK, n - p);
                LargeArrayUtils.arraycopy(a, offa + p, tmp, 0, len);
                MemorySegment.copy(tmp.getData(), 0, seg, JAVA_DOUBLE, 8L * p, (int) len);
            }
            plan.exec(op, seg, n, 0, scale);
            for (long p = 0; p < n; p += CHUNK) {
                final long len = Math.min((long) CHUNK, n - p);
                MemorySegment.copy(seg, JAVA_DOUBLE, 8L * p, tmp.getData(), 0, (int) len);
                LargeArrayUtils.arraycopy(tmp, 0, a, offa + p, len);
            }
        }
    }

    public static void execLarge(Plan plan, int op, FloatLargeArray a, long offa, boolean scale) {
        if (a.isConstant()) throw new IllegalArgumentException("The data array is constant.");
        if (!a.isLarge() && offa < Integer.MAX_VALUE) { plan.exec(op, a.getData(), offa, scale); return; }
        final long n = plan.elements(op);
        if (offa < 0 || offa + n > a.length()) throw new ArrayIndexOutOfBoundsException("array too short");
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment seg = arena.allocate(4L * n, 64);
            FloatLargeArray tmp = new FloatLargeArray((long) CHUNK, false);
            for (long p = 0; p < n; p += CHUNK) {
                final long len = Math.min((long) CHUNK, n - p);
                LargeArrayUtils.arraycopy(a, offa + p, tmp, 0, len);
                MemorySegment.copy(tmp.getData(), 0, seg, JAVA_FLOAT, 4L * p, (int) len);
            }
            plan.exec(op, seg, n, 0, scale);
            for (long p = 0; p < n; p += CHUNK) {
                final long len = Math.min((long) CHUNK, n - p);
                MemorySegment.copy(seg, JAVA_FLOAT, 4L * p, tmp.getData(), 0, (int) len);
                LargeArrayUtils.arraycopy(tmp, 0, a, offa + p, len);
            }
        }
    }

    // ------------------------------------------------------------------------------------------------ jagged arrays
    // double[rows][columns] / double[slices][rows][columns] (fft/DoubleFFT_2D.java:345, fft/DoubleFFT_3D.java:...): the
    // rows are gathered into one native block, transformed there and scattered back.  `full` (realForwardFull /
    // realInverseFull): every row holds 2*columns elements; the input occupies the first `columns` of each row in the
    // jagged form but is dense (rows*columns) in the flat form the library takes (fft/DoubleFFT_2D.java:1022-1075).
    public static void execJagged(Plan plan, int op, double[][] a, long rows, long columns, boolean full, boolean scale) {
        final long n = plan.elements(op);
        final int w = (int) (full ? 2 * columns : (n / rows));     // elements per row of the result (complex ops: 2*columns)
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment seg = arena.allocate(8L * n, 64);
            final int win = full ? (int) columns : w;
            for (int r = 0; r < rows; ++r) MemorySegment.copy(a[r], 0, seg, JAVA_DOUBLE, 8L * r * win, win);
            plan.exec(op, seg, n, 0, scale);
            for (int r = 0; r < rows; ++r) MemorySegment.copy(seg, JAVA_DOUBLE, 8L * r * w, a[r], 0, w);
        }
    }

    public static void execJagged(Plan plan, int op, float[][] a, long rows, long columns, boolean full, boolean scale) {
        final long n = plan.elements(op);
        final int w = (int) (full ? 2 * columns : (n / rows));
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment seg = arena.allocate(4L * n, 64);
            final int win = full ? (int) columns : w;
            for (int r = 0; r < rows; ++r) MemorySegment.copy(a[r], 0, seg, JAVA_FLOAT, 4L * r * win, win);
            plan.exec(op, seg, n, 0, scale);
            for (int r = 0; r < rows; ++r) MemorySegment.copy(seg, JAVA_FLOAT, 4L * r * w, a[r], 0, w);
        }
    }

    public static void execJagged(Plan plan, int op, double[][][] a, long slices, long rows, long columns, boolean full, boolean scale) {
        final long n = plan.elements(op);
        final int w = (int) (full ? 2 * columns : (n / (slices * rows)));
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment seg = arena.allocate(8L * n, 64);
            final int win = full ? (int) columns : w;
            for (int s = 0; s < slices; ++s)
                for (int r = 0; r < rows; ++r) MemorySegment.copy(a[s][r], 0, seg, JAVA_DOUBLE, 8L * (s * rows + r) * win, win);
            plan.exec(op, seg, n, 0, scale);
            for (int s = 0; s < slices; ++s)
                for (int r = 0; r < rows; ++r) MemorySegment.copy(seg, JAVA_DOUBLE, 8L * (s * rows + r) * w, a[s][r], 0, w);
        }
    }

    public static void execJagged(Plan plan, int op, float[][][] a, long slices, long rows, long columns, boolean full, boolean scale) {
        final long n = plan.elements(op);
        final int w = (int) (full ? 2 * columns : (n / (slices * rows)));
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment seg = arena.allocate(4L * n, 64);
            final int win = full ? (int) columns : w;
            for (int s = 0; s < slices; ++s)
                for (int r = 0; r < rows; ++r) MemorySegment.copy(a[s][r], 0, seg, JAVA_FLOAT, 4L * (s * rows + r) * win, win);
            plan.exec(op, seg, n, 0, scale);
            for (int s = 0; s < slices; ++s)
                for (int r = 0; r < rows; ++r) MemorySegment.copy(seg, JAVA_FLOAT, 4L * (s * rows + r) * w, a[s][r], 0, w);
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
