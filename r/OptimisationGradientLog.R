OptimisationGradientLog <- function(logHyp,trainingData,trainingTargets,trainingEvents,meanFuncForm,dimension,extraParam,covFuncForm,nSamples,noiseCorr,logHypNoiseCorrection,imposePriors) {
  #---------------------------------------------------------------------------------------------#
  # Drop-in for OptimisationGradientLog.R: k x 1 gradient in the order of logHyp (gps_nlml_grad).
  # Reuses the factorisation of the preceding ForOptimisationLog call when logHyp is unchanged.
  # ARD uses the analytic contraction (the reference's ARD branch, :58-69, is not the derivative
  # of its own likelihood).  NOT RUN IN THE BUILD IMAGE.
  #---------------------------------------------------------------------------------------------#
  if(imposePriors) stop('imposePriors=TRUE cannot run in the reference as shipped')
  h <- if(isTRUE(.gps$inOptim)) .gps$handle else
       GpsBind(trainingData,trainingTargets,trainingEvents,covFuncForm,meanFuncForm,noiseCorr,logHypNoiseCorrection,extraParam)
  # A non-positive-definite Ky is repaired by the library in BOTH fn and gr (same nearPD matrix, same value), so the two
  # agree on every point optim can accept; where no repair is possible fn has returned +Inf (see ForOptimisationLog.R)
  # and optim never asks for a gradient there.  If it happens anyway, say what is going on instead of a bare status.
  gradients <- tryCatch(.Call('gps_r_nlml_grad', h, as.double(logHyp)),
    error = function(e){
      if(grepl('^gps_not_pd', conditionMessage(e)))
        stop('OptimisationGradientLog: covariance matrix is not positive definite at a point the objective accepted, and it could not be repaired (n > 4096 or nearPD did not converge)')
      stop(e)
    })
  return(gradients)
}
