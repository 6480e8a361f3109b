ForOptimisationLog <- function(logHyp,trainingData,trainingTargets,trainingEvents,meanFuncForm,dimension,extraParam,covFuncForm,nSamples,noiseCorr,logHypNoiseCorrection,imposePriors){
  #-------------------------------------------------------------------------------------------------------------------#
  # Drop-in for ForOptimisationLog.R (same 12 arguments, returns the 1x1 negative log marginal likelihood).
  # Covariance, Cholesky, solves and the likelihood run in libgpsurv.so (gps_nlml).  NOT RUN IN THE BUILD IMAGE.
  #-------------------------------------------------------------------------------------------------------------------#
  if(imposePriors) stop('imposePriors=TRUE cannot run in the reference as shipped (LogPriorX takes 3 arguments, is called with 7)')
  h <- GpsBind(trainingData,trainingTargets,trainingEvents,covFuncForm,meanFuncForm,noiseCorr,logHypNoiseCorrection,extraParam)
  logMarLik <- tryCatch(.Call('gps_r_nlml', h, as.double(logHyp)),
    error = function(e){
      if(!grepl('^gps_not_pd', conditionMessage(e))) stop(e)
      # Reference fallback (ForOptimisationLog.R:43-51): nearest positive-definite matrix.  Rare, so it stays on the
      # host in R (Matrix::nearPD = repeated symmetric eigen-decompositions); the covariance still comes from the GPU.
      GpsNlmlNearPD(logHyp,trainingData,trainingTargets,trainingEvents,meanFuncForm,dimension,extraParam,covFuncForm,
                    nSamples,noiseCorr,logHypNoiseCorrection)
    })
  return(logMarLik)
}
