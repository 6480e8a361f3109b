GPPredict <- function(regressionStructure,parameterStructure,trainingTestStructure,logHypNoiseCorrection){
  #-------------------------------------------------------------------------------------------------------------------#
  # Drop-in for GPPredict.R: same arguments, returns list(funcMeanPred, varFunction, varData).
  # GPS2/GPS3 rebuild the factor from the passed noise correction and the current targets (:36-46);
  # GPR/GPS1 use the factor left resident by GPRegression (reuse_model).  NOT RUN IN THE BUILD IMAGE.
  #-------------------------------------------------------------------------------------------------------------------#
  meanFuncForm <- parameterStructure$meanFuncForm
  covFuncForm  <- parameterStructure$covFuncForm
  noiseCorr    <- parameterStructure$noiseCorr
  logHyp       <- regressionStructure$logHypChosen
  dimension    <- dim(trainingTestStructure$trainingData)[2]
  par <- c(as.double(logHyp$noise), as.double(logHyp$func), as.double(logHyp$length))
  if(meanFuncForm=='Linear') par <- c(par, as.double(logHyp$mean))
  corr <- (noiseCorr=='noiseCorrVec'|noiseCorr=='noiseCorrLearned')
  if(noiseCorr=='noiseCorrLearned') par <- c(par, as.double(logHypNoiseCorrection))
  testData <- as.matrix(trainingTestStructure$testData); storage.mode(testData) <- 'double'
  if(corr){
    h <- GpsBind(trainingTestStructure$trainingData,trainingTestStructure$trainingTargets,trainingTestStructure$events,
                 covFuncForm,meanFuncForm,noiseCorr,logHypNoiseCorrection,parameterStructure$extraParam)
    toReturn <- .Call('gps_r_predict', h, par, FALSE, testData)
  } else {
    # GPR / GPS1 predict from regressionStructure$L / alpha (GPPredict.R:12-13).  Those live on the GPU as the model the
    # finalize call of THAT GPRegression left resident; a regressionStructure from an earlier fit must not silently pick
    # up whatever model is resident now.
    if(is.null(regressionStructure$gpsToken) || is.null(.gps$handle) ||
       !identical(as.double(regressionStructure$gpsToken), as.double(.Call('gps_r_model_token', .gps$handle))))
      stop('GPPredict: the model resident on the GPU is not the one this regressionStructure came from; call GPRegression again (or predict before fitting another model)')
    toReturn <- .Call('gps_r_predict', .gps$handle, par, TRUE, testData)
  }
  return(toReturn)
}
