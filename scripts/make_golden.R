#!/usr/bin/env Rscript
# Pins for the CPU oracle (oracle/) from the REFERENCE ITSELF.  Run where R and the reference package are
# installed (e.g. a GPU box with baseline/_ref as an R library):
#     Rscript scripts/make_golden.R [lib.loc] > tests/golden/r_golden.json
# tests/test_r_golden.py loads the file when it exists and compares the oracle (and through it the CUDA
# path) with every value.  Written without jsonlite: base R only.  The reference pins R 3.4.0
# (packrat/packrat.lock:3); under R >= 3.6 sample() changed and under R >= 4.2 several `if (<data.frame>)`
# conditions of the reference stop -- the R version string is recorded so that the test can tell.
args = commandArgs(trailingOnly = TRUE)
if (length(args) >= 1) .libPaths(c(args[1], .libPaths()))
suppressMessages(library(soundgen))
num = function(x) paste0('[', paste(formatC(as.numeric(x), digits = 17, format = 'g'), collapse = ', '), ']')
out = list()
put = function(name, x) out[[name]] <<- num(x)

# --- R's random stream (oracle/rrng.py, csrc/rrng.h)
set.seed(1); put('runif_seed1', runif(5))
set.seed(1); put('rnorm_seed1', rnorm(6))
set.seed(42); put('rnorm_seed42', rnorm(5))
set.seed(1); put('rexp_seed1', rexp(3))
set.seed(7); put('rgamma_seed7_shape4_rate2', rgamma(8, 4, 2))
set.seed(7); put('rgamma_seed7_shape0.5_rate1', rgamma(8, 0.5, 1))
set.seed(7); put('rgamma_seed7_shape1600_rate40', rgamma(8, 1600, 40))
set.seed(3); put('rbinom_seed3_size1_p0.3', rbinom(10, 1, 0.3))
set.seed(1); put('sample10_seed1', sample(1:10))
set.seed(5); put('sample_prob_seed5', sapply(1:20, function(i) sample(1:3, 1, prob = c(.9, .05, .05))))

# --- SURVEY.md Appendix B known answers (hand-derived there; here from R)
put('getGlottalCycles', soundgen:::getGlottalCycles(seq(150, 200, length.out = 350), samplingRate = 3500))
up = soundgen:::upsample(c(100, 150, 130), 16000)
put('upsample_gc', up$gc); put('upsample_pitch', up$pitch)
put('clumper_2', soundgen:::clumper(c(1, 3, 2, 2, 2, 0, 0, 4, 4, 1, 1, 1, 1, 1, 3, 3), 2))
put('clumper_3', soundgen:::clumper(c(1, 3, 2, 2, 2, 0, 0, 4, 4, 1, 1, 1, 1, 1, 3, 3), 3))
put('addVectors_5', soundgen:::addVectors(1:6, rep(100, 3), insertionPoint = 5))
put('addVectors_m4', soundgen:::addVectors(1:6, rep(100, 3), insertionPoint = -4))
put('matchLengths_5', soundgen:::matchLengths(c(1, 2, 3), len = 5))
put('getRolloff', as.numeric(getRolloff(pitch_per_gc = c(100, 150, 130), nHarmonics = 20, rolloff = -12,
                                        rolloffOct = -2, rolloffKHz = -6)))

# --- loess contours (oracle/rloess.py, csrc/rloess.cuh): the default pitch anchors and a few more
ct = function(t, v, len, ...) suppressWarnings(soundgen:::getSmoothContour(anchors = data.frame(time = t, value = v), len = len, ...))
put('contour_default_pitch_1050', ct(c(0, .1, .9, 1), c(100, 150, 135, 100), 1050, samplingRate = 3500,
                                     valueFloor = 50, valueCeiling = 3500, thisIsPitch = TRUE))
put('contour_default_pitch_3500', ct(c(0, .1, .9, 1), c(100, 150, 135, 100), 3500, samplingRate = 3500,
                                     valueFloor = 50, valueCeiling = 3500, thisIsPitch = TRUE))
put('contour_3_anchors', ct(c(0, .38, 1), c(147, 163, 150), 875, samplingRate = 3500, valueFloor = 50,
                            valueCeiling = 3500, thisIsPitch = TRUE))
put('contour_6_anchors', ct(c(0, .05, .18, .45, .91, 1), c(221, 322, 346, 304, 273, 253), 11025,
                            samplingRate = 3500, valueFloor = 50, valueCeiling = 3500, thisIsPitch = TRUE))
put('contour_mouth', ct(c(0, .12, .86, 1), c(0, .52, .57, 0), 64, valueFloor = 0, valueCeiling = 1))
put('contour_noise_4', ct(c(-36, 8, 242, 333), c(-86, -24, -34, -118), 5904, valueFloor = -120, valueCeiling = 40,
                          samplingRate = 16000))

# --- whole calls: BASELINE config 0 and a few presets, each under its own seed
calls = list(
  cfg0_seed1 = list(seed = 1, call = quote(soundgen(sylLen = 1000))),
  cfg0_seed2 = list(seed = 2, call = quote(soundgen(sylLen = 1000))),
  t0_two_anchors = list(seed = 1, call = quote(soundgen(sylLen = 1000, pitchAnchors = c(100, 150), temperature = 0,
                                                        addSilence = 100))),
  preset_M1_Roar = list(seed = 2, call = parse(text = presets$M1$Roar)[[1]]),
  preset_Cat_Heat = list(seed = 22, call = parse(text = presets$Cat$Heat)[[1]]),
  preset_Misc_Seagull = list(seed = 32, call = parse(text = presets$Misc$Seagull)[[1]]))
for (nm in names(calls)) {
  set.seed(calls[[nm]]$seed)
  y = try(suppressWarnings(eval(calls[[nm]]$call)), silent = TRUE)
  if (!inherits(y, 'try-error')) {
    put(paste0('wave_', nm), y)
    put(paste0('stream_after_', nm), runif(2))      # where the call left R's stream
  }
}

cat('{\n  "R.version.string": "', R.version.string, '",\n  "soundgen.version": "',
    as.character(packageVersion('soundgen')), '",\n', sep = '')
cat(paste0('  "', names(out), '": ', unlist(out), collapse = ',\n'), '\n}\n')
