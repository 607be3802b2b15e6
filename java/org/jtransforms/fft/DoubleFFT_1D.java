/*
 * Drop-in org.jtransforms.fft.DoubleFFT_1D over libjtb200 (SOURCE ONLY -- see java/org/jtransforms/b200/Jtb200.java).
 * Same public surface as the reference class (fft/DoubleFFT_1D.java:53, :117, :204-1209); the plan tables, Bluestein
 * chirps and all arithmetic live on the GPU.  FloatFFT_1D is identical with float[] / Jtb200.F32.
 */
package org.jtransforms.fft;

import org.jtransforms.b200.Jtb200;
import pl.edu.icm.jlargearrays.DoubleLargeArray;

public final class DoubleFFT_1D {
    private final long n;
    private final Jtb200.Plan plan;

    public DoubleFFT_1D(long n) {                      // IllegalArgumentException("n must be greater than 0") as :119-121
        this.plan = new Jtb200.Plan(Jtb200.FFT, Jtb200.F64, n);
        this.n = n;
    }

    public void complexForward(double[] a) { complexForward(a, 0); }
    public void complexForward(double[] a, int offa) { plan.exec(Jtb200.C2C_FORWARD, a, offa, false); }
    public void complexForward(DoubleLargeArray a) { complexForward(a, 0); }
    public void complexForward(DoubleLargeArray a, long offa) { execLarge(Jtb200.C2C_FORWARD, a, offa, false); }

    public void complexInverse(double[] a, boolean scale) { complexInverse(a, 0, scale); }
    public void complexInverse(double[] a, int offa, boolean scale) { plan.exec(Jtb200.C2C_INVERSE, a, offa, scale); }
    public void complexInverse(DoubleLargeArray a, boolean scale) { execLarge(Jtb200.C2C_INVERSE, a, 0, scale); }
    public void complexInverse(DoubleLargeArray a, long offa, boolean scale) { execLarge(Jtb200.C2C_INVERSE, a, offa, scale); }

    public void realForward(double[] a) { realForward(a, 0); }
    public void realForward(double[] a, int offa) { plan.exec(Jtb200.R2C_PACKED, a, offa, false); }
    public void realForward(DoubleLargeArray a, long offa) { execLarge(Jtb200.R2C_PACKED, a, offa, false); }

    public void realForwardFull(double[] a) { realForwardFull(a, 0); }
    public void realForwardFull(double[] a, int offa) { plan.exec(Jtb200.R2C_FULL, a, offa, false); }
    public void realForwardFull(DoubleLargeArray a, long offa) { execLarge(Jtb200.R2C_FULL, a, offa, false); }

    public void realInverse(double[] a, boolean scale) { realInverse(a, 0, scale); }
    public void realInverse(double[] a, int offa, boolean scale) { plan.exec(Jtb200.C2R_PACKED, a, offa, scale); }
    public void realInverse(DoubleLargeArray a, long offa, boolean scale) { execLarge(Jtb200.C2R_PACKED, a, offa, scale); }

    public void realInverseFull(double[] a, boolean scale) { realInverseFull(a, 0, scale); }
    public void realInverseFull(double[] a, int offa, boolean scale) { plan.exec(Jtb200.C2R_FULL, a, offa, scale); }
    public void realInverseFull(DoubleLargeArray a, long offa, boolean scale) { execLarge(Jtb200.C2R_FULL, a, offa, scale); }

    /** DoubleLargeArray: heap-backed -> its double[]; off-heap (isLarge) -> native address, no 2^31 limit. */
    private void execLarge(int op, DoubleLargeArray a, long offa, boolean scale) {
        if (!a.isLarge() && !a.isConstant()) { plan.exec(op, a.getData(), offa, scale); return; }
        if (a.isConstant()) throw new IllegalArgumentException("The data array is constant.");
        plan.exec(op, java.lang.foreign.MemorySegment.ofAddress(a.nativePointer()).reinterpret(a.length() * 8L), offa, scale);
    }
}
