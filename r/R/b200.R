# R shells over libsoundgen_b200 (see INTEGRATION.md).  Signatures are those of the reference
# (R/source.R:173-205, R/sourceSpectrum.R:71-82, :261-283); bodies only marshal arguments and
# keep R's RNG stream where the reference would leave it.
#' @useDynLib soundgen, .registration = TRUE

# Draws `cap` normals, lets the device consume what the reference would have drawn, then puts
# R's stream exactly where generateHarmonics() would have left it (SURVEY.md 8a, RNG ledger).
.with_normals = function(cap, f) {
  if (!exists('.Random.seed', envir = globalenv())) runif(1)
  seed = get('.Random.seed', envir = globalenv())
  z = rnorm(cap)
  res = f(z)
  assign('.Random.seed', seed, envir = globalenv())
  if (res$z_used > 0) invisible(rnorm(res$z_used))
  res
}

generateHarmonics = function(pitch, attackLen = 50, nonlinBalance = 0, nonlinDep = 0, jitterDep = 0,
                             jitterLen = 1, vibratoFreq = 100, vibratoDep = 0, shimmerDep = 0,
                             creakyBreathy = 0, rolloff = -18, rolloffOct = -2, rolloffKHz = -6,
                             rolloffParab = 0, rolloffParabHarm = 3, rolloffLip = 6, rolloff_perAmpl = 12,
                             temperature = 0, pitchDriftDep = .5, pitchDriftFreq = .125,
                             randomWalk_trendStrength = .5, shortestEpoch = 300, subFreq = 100, subDep = 0,
                             amDep = 0, amFreq = 30, amplAnchors = NA, overlap = 75, samplingRate = 16000,
                             pitchFloor = 75, pitchCeiling = 3500, pitchSamplingRate = 3500,
                             throwaway = -120) {
  pars = list(attackLen = attackLen, nonlinBalance = nonlinBalance, jitterDep = jitterDep,
              jitterLen = jitterLen, vibratoFreq = vibratoFreq, vibratoDep = vibratoDep,
              shimmerDep = shimmerDep, rolloff = rolloff, rolloffOct = rolloffOct, rolloffKHz = rolloffKHz,
              rolloffParab = rolloffParab, rolloffParabHarm = rolloffParabHarm,
              rolloff_perAmpl = rolloff_perAmpl, temperature = temperature, pitchDriftDep = pitchDriftDep,
              pitchDriftFreq = pitchDriftFreq, randomWalk_trendStrength = randomWalk_trendStrength,
              shortestEpoch = shortestEpoch, subFreq = subFreq, subDep = subDep, samplingRate = samplingRate,
              pitchFloor = pitchFloor, pitchCeiling = pitchCeiling, pitchSamplingRate = pitchSamplingRate,
              throwaway = throwaway)
  ampl = NULL
  if (is.list(amplAnchors)) {
    # 1, 2 or > 10 anchors are evaluated on the device; 3-10 anchors (loess) are not supported
    # by this entry point: use the staged interface described in INTEGRATION.md
    if (nrow(amplAnchors) > 2 && nrow(amplAnchors) <= 10)
      stop('amplAnchors with 3-10 anchors need the staged (loess on host) interface')
    ampl = cbind(amplAnchors$time, amplAnchors$value)
  }
  cap = 2 * length(pitch) + 64      # rw (<= nGC) + jitter idx (<= nGC) + drift (<= nGC) + shimmer (nGC)
  res = .with_normals(cap, function(z)
    .Call(sg_generate_harmonics, as.numeric(pitch), pars, z, ampl))
  res$waveform
}

getRolloff = function(pitch_per_gc = c(440), nHarmonics = 100, rolloff = -12, rolloffOct = -2,
                      rolloffParab = 0, rolloffParabHarm = 2, rolloffParabCeiling = NULL, rolloffKHz = -6,
                      baseline = 200, throwaway = -120, samplingRate = 16000, plot = FALSE) {
  r = .Call(sg_get_rolloff, as.numeric(pitch_per_gc), as.integer(nHarmonics), as.numeric(rolloff),
            as.numeric(rolloffOct), as.numeric(rolloffKHz), rolloffParab, rolloffParabHarm,
            rolloffParabCeiling, baseline, throwaway, samplingRate)
  # plotting (sourceSpectrum.R:149-176) stays in R and is unchanged
  r
}

# generateNoise (R/source.R:57-138): the breathing-strength contour with 3-10 anchors uses loess, which
# stays in R (getSmoothContour is unchanged); 1, 2 or > 10 anchors are evaluated on the device.  The
# runif() draws are made here, in the reference's order and number (source.R:88-111).
generateNoise = function(len, noiseAnchors = data.frame(time = c(0, 300), value = c(-120, -120)),
                         rolloffNoise = -6, attackLen = 10, windowLength_points = 1024,
                         samplingRate = 16000, overlap = 75, throwaway = -120, filterNoise = NA) {
  anchors = NULL
  strength = NULL
  n = nrow(noiseAnchors)
  if (n > 2 && n <= 10) {
    strength = getSmoothContour(len = len, anchors = noiseAnchors,
                                valueFloor = permittedValues['noiseAmpl', 'low'],
                                valueCeiling = permittedValues['noiseAmpl', 'high'],
                                samplingRate = samplingRate, plot = FALSE)
  } else {
    anchors = cbind(noiseAnchors$time, noiseAnchors$value)
  }
  step = seq(1, len + windowLength_points,
             by = windowLength_points - (overlap * windowLength_points / 100))
  nr = windowLength_points / 2
  u = runif(nr * length(step))
  filt = NULL
  if (is.matrix(filterNoise) || (is.numeric(filterNoise) && length(filterNoise) > 1)) filt = as.matrix(filterNoise)
  pars = list(rolloffNoise = rolloffNoise, attackLen = attackLen, windowLength_points = windowLength_points,
              samplingRate = samplingRate, overlap = overlap, throwaway = throwaway)
  .Call(sg_generate_noise, as.integer(len), anchors, strength, u, filt, pars)
}
