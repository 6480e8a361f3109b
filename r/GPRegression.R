GPRegression <- function(logHyp,parameterStructure,trainingTestStructure,logHypNoiseCorrection){
  #-------------------------------------------------------------------------------------------------------------------#
  # Drop-in for GPRegression.R: same arguments, same returned list.  stats::optim keeps driving
  # ForOptimisationLog / OptimisationGradientLog (now GPU-backed); the post-optim block (:60-71) is one
  # call (gps_regress_finalize) that also leaves the factor resident for GPPredict.  NOT RUN IN THE BUILD IMAGE.
  #-------------------------------------------------------------------------------------------------------------------#
  trainingData    <- trainingTestStructure$trainingData
  trainingTargets <- trainingTestStructure$trainingTargets
  trainingEvents  <- trainingTestStructure$events
  nTraining       <- trainingTestStructure$nTraining
  dimension       <- trainingTestStructure$dimension
  noiseCorr       <- parameterStructure$noiseCorr
  modelType       <- parameterStructure$modelType
  optimType       <- parameterStructure$optimType
  if(modelType=='survival') maxit <- parameterStructure$maxitSurvival else maxit <- parameterStructure$maxit
  meanFuncForm    <- parameterStructure$meanFuncForm
  covFuncForm     <- parameterStructure$covFuncForm
  extraParam      <- parameterStructure$extraParam
  imposePriors    <- parameterStructure$imposePriors

  logHypVec <- as.matrix(unlist(logHyp))
  if(meanFuncForm=='Zero') logHypVec <- logHypVec[-c((length(logHypVec)-dimension):length(logHypVec)),,drop=FALSE]
  if(noiseCorr=='noiseCorrLearned') logHypVec <- rbind(logHypVec,logHypNoiseCorrection)

  # bind ONCE: X, targets, events, options and the GPS3 noise correction are constant over this optim() run; fn / gr skip
  # the per-call O(n d) identity checks while .gps$inOptim is set (reset on exit, also on error)
  GpsBind(trainingData,trainingTargets,trainingEvents,covFuncForm,meanFuncForm,noiseCorr,logHypNoiseCorrection,extraParam)
  .gps$inOptim <- TRUE; .gps$optimType <- optimType
  on.exit({.gps$inOptim <- FALSE; .gps$optimType <- NULL}, add=TRUE)
  optimOutput <- optim(logHypVec,fn=ForOptimisationLog,gr=OptimisationGradientLog,trainingData=trainingData,trainingTargets=trainingTargets,
                       trainingEvents=trainingEvents,meanFuncForm=meanFuncForm,dimension=dimension,extraParam=extraParam,covFuncForm=covFuncForm,
                       nSamples=nTraining,noiseCorr=noiseCorr,logHypNoiseCorrection=logHypNoiseCorrection,imposePriors=imposePriors,
                       method=optimType,control=list('maxit'=maxit))
  .gps$inOptim <- FALSE
  parOpt    <- as.double(optimOutput$par)
  objective <- optimOutput$value

  h   <- GpsBind(trainingData,trainingTargets,trainingEvents,covFuncForm,meanFuncForm,noiseCorr,logHypNoiseCorrection,extraParam)
  fin <- .Call('gps_r_finalize', h, parOpt, as.integer(nTraining), TRUE)

  logHypVec <- as.matrix(parOpt)
  if(noiseCorr=='noiseCorrLearned'){
    logHypNoiseCorrection <- logHypVec[length(logHypVec),,drop=FALSE]
    logHypVec             <- logHypVec[-length(logHypVec),,drop=FALSE]
  }
  nLength <- if(covFuncForm=='ARD') dimension else 1
  switch(meanFuncForm,
    'Linear' = {meanHyp <- logHypVec[(length(logHypVec)-dimension):length(logHypVec),,drop=FALSE]},
    'Zero'   = {meanHyp <- matrix(rep(0,dimension+1),nrow=dimension+1)})
  logHypChosen <- list('noise'=log(exp(logHypVec[1,,drop=FALSE])),'func'=log(exp(logHypVec[2,,drop=FALSE])),
                       'length'=log(exp(logHypVec[3:(2+nLength),,drop=FALSE])),'mean'=meanHyp)
  toReturn <- list('logHyp'=logHyp,'meanTraining'=fin$meanTraining,'K'=fin$K,'L'=fin$L,'alpha'=fin$alpha,
                   'logHypChosen'=logHypChosen,'logMarLik'=fin$logMarLik,'objective'=objective,
                   'gpsToken'=fin$gpsToken)     # identity of the model left resident on the GPU (checked by GPPredict)
  if(noiseCorr=='noiseCorrVec'|noiseCorr=='noiseCorrLearned'){
    if(logHypNoiseCorrection[1]==-Inf) logHypNoiseCorrection <- -10^10
    toReturn$logHypNoiseCorrection <- logHypNoiseCorrection
  }
  return(toReturn)
}
