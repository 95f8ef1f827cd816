package org.apache.spark.ml.commons

import breeze.linalg.{DenseMatrix => BDM, DenseVector => BDV}
import org.apache.spark.ml.commons.kernel._
import org.apache.spark.ml.linalg.Vector
import org.apache.spark.rdd.RDD

/** JNI binding of libsgp.so (include/sgp.h).  One native context per executor task; device = partitionId % nGPUs. */
private[ml] object NativeProjectedProcess {
  System.loadLibrary("sgp_jni")
  @native def create(device: Int): Long
  @native def destroy(ctx: Long): Unit
  @native def begin(ctx: Long, types: Array[Int], scales: Array[Double], sigmas: Array[Double], betas: Array[Double],
                    activeSet: Array[Double], m: Int, d: Int): Unit
  @native def accumulate(ctx: Long, x: Array[Double], y: Array[Double], n: Long): Unit
  @native def finish(ctx: Long, g: Array[Double], b: Array[Double]): Unit
  @native def magic(ctx: Long, g: Array[Double], b: Array[Double], magicVector: Array[Double],
                    magicMatrix: Array[Double]): Unit

  /** Flattens the kernel DSL tree into (type, scale, sigma, beta) terms; `scale` multiplies down the tree
    * (ScalarTimesKernel.scala:20-28), Eye terms are kept (they carry whiteNoiseVar / the K_mm diagonal). */
  def flatten(k: Kernel, scale: Double = 1d): Seq[(Int, Double, Double, Array[Double])] = k match {
    case s: SumOfKernels       => flatten(s.kernel1, scale) ++ flatten(s.kernel2, scale)   // needs the two vals exposed
    case c: ScalarTimesKernel  => flatten(c.innerKernel, scale * c.scalar)                 // ditto (kernel, C)
    case a: ARDRBFKernel       => Seq((0, scale, 0d, a.getHyperparameters.toArray))
    case r: RBFKernel          => Seq((1, scale, r.getHyperparameters(0), Array.empty[Double]))
    case _: EyeKernel          => Seq((2, scale, 0d, Array.empty[Double]))
  }
}

/** Drop-in for ProjectedGaussianProcessHelper (commons/ProjectedGaussianProcessHelper.scala): same two methods,
  * same return types; mix this trait into GaussianProcessCommons instead of the original. */
private[ml] trait NativeProjectedGaussianProcessHelper extends ProjectedGaussianProcessHelper {
  import NativeProjectedProcess._

  def nGPUs: Int = 8

  override def getMatrixKmnKnmAndVectorKmny(expertLabelsAndKernels: RDD[(BDV[Double], Kernel)],
                                            activeSet: Array[Vector]): (BDM[Double], BDV[Double]) = {
    val m = activeSet.length
    val d = activeSet.head.size
    val z = activeSet.flatMap(_.toArray)
    val terms = flatten(expertLabelsAndKernels.first()._2)        // every expert carries the same kernel / theta
    val (ty, sc, sg, be) = (terms.map(_._1).toArray, terms.map(_._2).toArray, terms.map(_._3).toArray,
      terms.flatMap(_._4).toArray)
    // seqOp of PGPH:25-30 becomes: one native context per partition, all experts of the partition in one shard
    val partials = expertLabelsAndKernels.mapPartitionsWithIndex { (pid, it) =>
      val ctx = create(pid % nGPUs)
      try {
        begin(ctx, ty, sc, sg, be, z, m, d)
        it.grouped(256).foreach { experts =>                      // pack a few hundred experts per native call
          val x = experts.flatMap(_._2.getTrainingVectors.flatMap(_.toArray)).toArray
          val y = experts.flatMap(_._1.toArray).toArray
          accumulate(ctx, x, y, y.length.toLong)
        }
        val g = new Array[Double](m * m); val b = new Array[Double](m)
        finish(ctx, g, b)
        Iterator.single((new BDM(m, m, g), new BDV(b)))
      } finally destroy(ctx)
    }
    partials.treeReduce { case ((g1, b1), (g2, b2)) => (g1 += g2, b1 += b2) }   // combOp, PGPH:31-35
  }

  override def getMagicVector(kernel: Kernel, matrixKmnKnm: BDM[Double], vectorKmny: BDV[Double],
                              activeSet: Array[Vector], optimalHyperparameter: BDV[Double]) = {
    val m = activeSet.length
    val d = activeSet.head.size
    val terms = flatten(kernel)
    val ctx = create(0)
    try {
      begin(ctx, terms.map(_._1).toArray, terms.map(_._2).toArray, terms.map(_._3).toArray,
        terms.flatMap(_._4).toArray, activeSet.flatMap(_.toArray), m, d)
      val mv = new Array[Double](m); val mm = new Array[Double](m * m)
      magic(ctx, matrixKmnKnm.toArray, vectorKmny.toArray, mv, mm)   // throws NotPositiveDefiniteException like PGPH:62-65
      (new BDV(mv), new BDM(m, m, mm))
    } finally destroy(ctx)
  }
}
