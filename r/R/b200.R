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

# getSpectralEnvelope (R/sourceSpectrum.R:261-283): same signature.  The deterministic part runs on the
# device; with temperature > 0 the stochastic block (:346-415) is R code that stays here unchanged (it draws
# rgamma / rnorm from R's stream) and hands its formants_upsampled matrices down as tracks.
getSpectralEnvelope = function(nr, nc, formants = NA, formantDep = 1, rolloffLip = 6, mouthAnchors = NA,
                               mouthOpenThres = 0, openMouthBoost = 0, vocalTract = NULL, temperature = 0,
                               formDrift = .3, formDisp = .2, formantDepStoch = 30, smoothLinearFactor = 1,
                               samplingRate = 16000, speedSound = 35400, plot = FALSE, duration = NULL,
                               colorTheme = c('bw', 'seewave', '...')[1], nCols = 100, xlab = 'Time',
                               ylab = 'Frequency, kHz', ...) {
  if (class(formants)[1] == 'character') formants = convertStringToFormants(formants)
  tracks = FALSE
  fm = NULL; fn = NULL
  if (is.list(formants)) {
    formants = lapply(formants, as.data.frame)
    if (temperature > 0) {
      # The stochastic block stays R code on R's stream: the package keeps lines :321-415 of the original
      # function as an internal helper that returns `formants_upsampled` (one data.frame of nc rows per
      # formant, the drawn pseudo-formants appended); see INTEGRATION.md, "what stays in R".
      formants = .formants_upsampled_stochastic(formants, nc, temperature, formDrift, formDisp, formantDep,
                                                formantDepStoch, vocalTract, smoothLinearFactor, samplingRate,
                                                speedSound)
      tracks = TRUE
    }
    fm = do.call(rbind, lapply(formants, function(f) cbind(f$time, f$freq, f$amp, f$width)))
    fn = as.integer(sapply(formants, nrow))
  }
  mouth = NULL
  if (is.list(mouthAnchors) && !any(is.na(mouthAnchors))) mouth = cbind(mouthAnchors$time, mouthAnchors$value)
  pars = list(formantDep = formantDep, rolloffLip = rolloffLip, mouthOpenThres = mouthOpenThres,
              openMouthBoost = openMouthBoost, vocalTract = if (is.numeric(vocalTract)) vocalTract else NaN,
              samplingRate = samplingRate, speedSound = speedSound, smoothLinearFactor = smoothLinearFactor)
  m = .Call(sg_get_spectral_envelope, as.integer(nr), as.integer(nc), fm, fn, tracks, mouth, pars)
  # plotting (sourceSpectrum.R:543-563) stays in R and is unchanged
  m
}

# The filter block of soundgen() (R/soundgen.R:736-808) as one call: stft -> envelope -> istft -> / max.
filterSound_b200 = function(sound, spectralEnvelope, windowLength_points, overlap = 75) {
  if (sum(sound) == 0) return(sound)                                  # :736-739
  .Call(sg_filter, as.numeric(sound), spectralEnvelope, as.integer(windowLength_points), overlap)
}

.anchor_matrix = function(a, t_hi = 1) {
  if (is.numeric(a) && length(a) > 0) a = data.frame(time = seq(0, t_hi, length.out = length(a)), value = a)
  if (is.list(a) && !any(is.na(a))) cbind(as.numeric(a$time), as.numeric(a$value)) else NULL
}
.formant_matrices = function(f) {
  if (is.character(f)) f = convertStringToFormants(f)
  if (!is.list(f)) return(NULL)
  lapply(f, function(x) { x = as.data.frame(x); cbind(x$time, x$freq, x$amp, x$width) })
}

# soundgen() (R/soundgen.R:208-277), same arguments: the whole call in the library.  The host stage runs in
# the library's front-end on R's own random stream: .Random.seed goes in, the call draws what the reference
# would have drawn, in its order, and the advanced seed is put back.  play / savePath / plot stay R code.
soundgen = function(repeatBout = 1, nSyl = 1, sylLen = 300, pauseLen = 200,
                    pitchAnchors = data.frame(time = c(0, .1, .9, 1), value = c(100, 150, 135, 100)),
                    pitchAnchorsGlobal = NA, temperature = 0.025,
                    tempEffects = list(sylLenDep = .02, formDrift = .3, formDisp = .2, pitchDriftDep = .5,
                                       pitchDriftFreq = .125, pitchAnchorsDep = .05, noiseAnchorsDep = .1,
                                       amplAnchorsDep = .1),
                    maleFemale = 0, creakyBreathy = 0, nonlinBalance = 0, nonlinDep = 50, jitterLen = 1,
                    jitterDep = 3, vibratoFreq = 5, vibratoDep = 0, shimmerDep = 0, attackLen = 50,
                    rolloff = -12, rolloffOct = -12, rolloffKHz = -6, rolloffParab = 0, rolloffParabHarm = 3,
                    rolloffLip = 6,
                    formants = list(f1 = list(time = 0, freq = 860, amp = 30, width = 120),
                                    f2 = list(time = 0, freq = 1280, amp = 40, width = 120),
                                    f3 = list(time = 0, freq = 2900, amp = 25, width = 200)),
                    formantDep = 1, formantDepStoch = 30, vocalTract = 15.5, subFreq = 100, subDep = 100,
                    shortestEpoch = 300, amDep = 0, amFreq = 30, amShape = 0,
                    noiseAnchors = data.frame(time = c(0, 300), value = c(-120, -120)), formantsNoise = NA,
                    rolloffNoise = -14, mouthAnchors = data.frame(time = c(0, 1), value = c(.5, .5)),
                    amplAnchors = NA, amplAnchorsGlobal = NA, samplingRate = 16000, windowLength = 50,
                    overlap = 75, addSilence = 100, pitchFloor = 50, pitchCeiling = 3500,
                    pitchSamplingRate = 3500, throwaway = -120,
                    invalidArgAction = c('adjust', 'abort', 'ignore')[1], plot = FALSE, play = FALSE,
                    savePath = NA, ...) {
  if (RNGkind()[1] != 'Mersenne-Twister' || RNGkind()[2] != 'Inversion')
    stop('soundgen_b200 follows the Mersenne-Twister / Inversion stream (the defaults)')
  if (!exists('.Random.seed', envir = globalenv())) set.seed(NULL)
  seed = get('.Random.seed', envir = globalenv())
  num = c('repeatBout', 'nSyl', 'sylLen', 'pauseLen', 'temperature', 'maleFemale', 'creakyBreathy',
          'nonlinBalance', 'nonlinDep', 'jitterLen', 'jitterDep', 'vibratoFreq', 'vibratoDep', 'shimmerDep',
          'attackLen', 'rolloff', 'rolloffOct', 'rolloffKHz', 'rolloffParab', 'rolloffParabHarm', 'rolloffLip',
          'formantDep', 'formantDepStoch', 'vocalTract', 'subFreq', 'subDep', 'shortestEpoch', 'amDep',
          'amFreq', 'amShape', 'rolloffNoise', 'samplingRate', 'windowLength', 'overlap', 'addSilence',
          'pitchFloor', 'pitchCeiling', 'pitchSamplingRate', 'throwaway')
  args = c(lapply(setNames(num, num), function(n) as.numeric(get(n))), lapply(tempEffects, as.numeric))
  anchors = list(.anchor_matrix(pitchAnchors), .anchor_matrix(pitchAnchorsGlobal),
                 .anchor_matrix(noiseAnchors, sylLen), .anchor_matrix(mouthAnchors),
                 .anchor_matrix(amplAnchors), .anchor_matrix(amplAnchorsGlobal))
  opts = list(invalidArgAction = match(invalidArgAction, c('adjust', 'abort', 'ignore')) - 1,
              contour_method = 0,
              sample_rejection = as.numeric(getRversion() >= '3.6.0' && RNGkind()[3] == 'Rejection'))
  res = .Call(sg_soundgen, args, anchors, .formant_matrices(formants), .formant_matrices(formantsNoise),
              opts, seed)
  assign('.Random.seed', res$seed, envir = globalenv())
  if (nchar(res$warnings) > 0) for (w in strsplit(res$warnings, '\n')[[1]]) warning(w)
  if (res$status != 0) stop(paste('soundgen_b200: call failed with status', res$status))
  bout = res$waveform
  if (play) playme(bout, samplingRate = samplingRate)                 # R/soundgen.R:851-860, unchanged
  if (!is.na(savePath)) seewave::savewav(bout, filename = savePath, f = samplingRate)
  if (plot) spectrogram(bout, samplingRate = samplingRate, ...)
  return(bout)
}

# The batched entry the reference lacks: a list of soundgen() argument lists and one seed per call; call i
# runs as `set.seed(seeds[i]); do.call(soundgen, calls[[i]])` would, but all calls cross the device as one
# batch.  Returns a list of waveforms (NULL where the reference would have stopped).
soundgen_batch = function(calls, seeds = seq_along(calls), invalidArgAction = 'adjust') {
  d = formals(soundgen)
  packed = lapply(calls, function(cl) {
    g = function(n) if (!is.null(cl[[n]])) cl[[n]] else eval(d[[n]])
    num = names(d)[sapply(names(d), function(n) is.numeric(eval(d[[n]])) && length(eval(d[[n]])) == 1)]
    te = modifyList(eval(d$tempEffects), if (is.null(cl$tempEffects)) list() else cl$tempEffects)
    list(c(lapply(setNames(num, num), function(n) as.numeric(g(n))), lapply(te, as.numeric)),
         list(.anchor_matrix(g('pitchAnchors')), .anchor_matrix(g('pitchAnchorsGlobal')),
              .anchor_matrix(g('noiseAnchors'), g('sylLen')), .anchor_matrix(g('mouthAnchors')),
              .anchor_matrix(g('amplAnchors')), .anchor_matrix(g('amplAnchorsGlobal'))),
         .formant_matrices(g('formants')), .formant_matrices(g('formantsNoise')))
  })
  opts = list(invalidArgAction = match(invalidArgAction, c('adjust', 'abort', 'ignore')) - 1, contour_method = 0,
              sample_rejection = as.numeric(getRversion() >= '3.6.0' && RNGkind()[3] == 'Rejection'))
  .Call(sg_soundgen_batch, packed, opts, as.integer(seeds))
}
