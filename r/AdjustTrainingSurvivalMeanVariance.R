AdjustTrainingSurvivalMeanVariance <- function(a,mu,sigma){
  #---------------------------------------------------------------------------------------------------------------------------------#
  # Drop-in for AdjustTrainingSurvivalMeanVariance.R; also accepts vectors (one launch for all censored rows).
  # NOT RUN IN THE BUILD IMAGE.
  #---------------------------------------------------------------------------------------------------------------------------------#
  if(is.null(.gps$handle)) .gps$handle <- .Call('gps_r_create', 0L)
  toReturn <- .Call('gps_r_adjust', .gps$handle, as.double(a), as.double(mu), as.double(sigma))
  return(toReturn)
}
