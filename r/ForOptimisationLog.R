ForOptimisationLog <- function(logHyp,trainingData,trainingTargets,trainingEvents,meanFuncForm,dimension,extraParam,covFuncForm,nSamples,noiseCorr,logHypNoiseCorrection,imposePriors){
  #-------------------------------------------------------------------------------------------------------------------#
  # Drop-in for ForOptimisationLog.R (same 12 arguments, returns the 1x1 negative log marginal likelihood).
  # Covariance, Cholesky, solves and the likelihood run in libgpsurv.so (gps_nlml).  NOT RUN IN THE BUILD IMAGE.
  #-------------------------------------------------------------------------------------------------------------------#
  if(imposePriors) stop('imposePriors=TRUE cannot run in the reference as shipped (LogPriorX takes 3 arguments, is called with 7)')
  # inside GPRegression's optim() the data were bound once (X, targets, options are constant over the run)
  h <- if(isTRUE(.gps$inOptim)) .gps$handle else
       GpsBind(trainingData,trainingTargets,trainingEvents,covFuncForm,meanFuncForm,noiseCorr,logHypNoiseCorrection,extraParam)
  logMarLik <- tryCatch(.Call('gps_r_nlml', h, as.double(logHyp)),
    error = function(e){
      if(!grepl('^gps_not_pd', conditionMessage(e))) stop(e)
      # The library repairs a non-positive-definite Ky itself (Matrix::nearPD's algorithm on the device, n <= 4096:
      # ForOptimisationLog.R:43-51); this branch is only reached beyond that size or when the repair did not converge.
      # fn and gr must agree on which points are acceptable: a gradient method could accept a host-repaired point and
      # then ask gr there, which has no host fallback -> under gradient methods the point is rejected (+Inf) instead.
      if(!is.null(.gps$optimType) && .gps$optimType %in% c('BFGS','CG','L-BFGS-B')) return(matrix(Inf,1,1))
      GpsNlmlNearPD(logHyp,trainingData,trainingTargets,trainingEvents,meanFuncForm,dimension,extraParam,covFuncForm,
                    nSamples,noiseCorr,logHypNoiseCorrection)
    })
  return(logMarLik)
}
