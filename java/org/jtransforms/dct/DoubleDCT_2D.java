/*
 * Drop-in org.jtransforms.dct.DoubleDCT_2D over libjtb200 (SOURCE ONLY).  Reference: dct/DoubleDCT_2D.java:75-98,
 * :104-183 (forward), :360-440 (inverse).  DoubleDST_2D / DoubleDHT_2D differ only in the plan kind (and DHT's
 * forward has no scale argument, dht/DoubleDHT_2D.java:102).
 */
package org.jtransforms.dct;

import org.jtransforms.b200.Jtb200;

public final class DoubleDCT_2D {
    private final Jtb200.Plan plan;

    public DoubleDCT_2D(long rows, long columns) {
        this.plan = new Jtb200.Plan(Jtb200.DCT, Jtb200.F64, rows, columns);
    }

    public void forward(double[] a, boolean scale) { plan.exec(Jtb200.R2R_FORWARD, a, 0, scale); }
    public void inverse(double[] a, boolean scale) { plan.exec(Jtb200.R2R_INVERSE, a, 0, scale); }
}
