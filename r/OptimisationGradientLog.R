OptimisationGradientLog <- function(logHyp,trainingData,trainingTargets,trainingEvents,meanFuncForm,dimension,extraParam,covFuncForm,nSamples,noiseCorr,logHypNoiseCorrection,imposePriors) {
  #---------------------------------------------------------------------------------------------#
  # Drop-in for OptimisationGradientLog.R: k x 1 gradient in the order of logHyp (gps_nlml_grad).
  # Reuses the factorisation of the preceding ForOptimisationLog call when logHyp is unchanged.
  # ARD uses the analytic contraction (the reference's ARD branch, :58-69, is not the derivative
  # of its own likelihood).  NOT RUN IN THE BUILD IMAGE.
  #---------------------------------------------------------------------------------------------#
  if(imposePriors) stop('imposePriors=TRUE cannot run in the reference as shipped')
  h <- GpsBind(trainingData,trainingTargets,trainingEvents,covFuncForm,meanFuncForm,noiseCorr,logHypNoiseCorrection,extraParam)
  gradients <- .Call('gps_r_nlml_grad', h, as.double(logHyp))
  return(gradients)
}
