CovFunc <-function(x1,x2,x,y,extraParam,varFuncSqHyp,lengthHyp,covFuncForm){
  #--------------------------------------------------------------------------#
  # Drop-in for CovFunc.R ('SqExp', 'ARD', 'Matern' with extraParam 1/3/5; 'RF', 'InformedARD' stop with an error).
  # NOT RUN IN THE BUILD IMAGE.
  #--------------------------------------------------------------------------#
  if(is.null(.gps$handle)) .gps$handle <- .Call('gps_r_create', 0L)
  x1 <- as.matrix(x1); x2 <- as.matrix(x2); storage.mode(x1) <- 'double'; storage.mode(x2) <- 'double'
  covariance <- .Call('gps_r_cov', .gps$handle, x1, if(identical(x1,x2)) NULL else x2, as.double(varFuncSqHyp),
                      as.double(lengthHyp), GpsCovCode(covFuncForm, extraParam), .gps$ardMode)
  return(covariance)
}
