/*
 * Drop-in org.jtransforms.fft.DoubleFFT_3D over libjtb200 (SOURCE ONLY).  Reference: fft/DoubleFFT_3D.java:88-124
 * (constructor), :145-325 (complexForward), :1339-1355 (realForward), :1629 (realInverse).  The plan is immutable,
 * so unlike the reference (which mutates its strides during a call, :149-162) instances are re-entrant.
 */
package org.jtransforms.fft;

import org.jtransforms.b200.Jtb200;

public final class DoubleFFT_3D {
    private final long slices, rows, columns;
    private final Jtb200.Plan plan;

    public DoubleFFT_3D(long slices, long rows, long columns) {   // "slices, rows and columns must be greater than 1"
        this.plan = new Jtb200.Plan(Jtb200.FFT, Jtb200.F64, slices, rows, columns);
        this.slices = slices; this.rows = rows; this.columns = columns;
    }

    public void complexForward(double[] a) { plan.exec(Jtb200.C2C_FORWARD, a, 0, false); }
    public void complexInverse(double[] a, boolean scale) { plan.exec(Jtb200.C2C_INVERSE, a, 0, scale); }
    public void realForward(double[] a) { plan.exec(Jtb200.R2C_PACKED, a, 0, false); }     // power-of-two sizes only
    public void realForwardFull(double[] a) { plan.exec(Jtb200.R2C_FULL, a, 0, false); }
    public void realInverse(double[] a, boolean scale) { plan.exec(Jtb200.C2R_PACKED, a, 0, scale); }
    public void realInverseFull(double[] a, boolean scale) { plan.exec(Jtb200.C2R_FULL, a, 0, scale); }

    /** double[][][] overload (fft/DoubleFFT_3D.java:346): rows are staged through one flat pinned buffer. */
    public void complexForward(double[][][] a) {
        int s = (int) slices, r = (int) rows, c2 = (int) (2 * columns);
        double[] flat = new double[s * r * c2];
        for (int i = 0; i < s; i++) for (int j = 0; j < r; j++) System.arraycopy(a[i][j], 0, flat, (i * r + j) * c2, c2);
        complexForward(flat);
        for (int i = 0; i < s; i++) for (int j = 0; j < r; j++) System.arraycopy(flat, (i * r + j) * c2, a[i][j], 0, c2);
    }
}
