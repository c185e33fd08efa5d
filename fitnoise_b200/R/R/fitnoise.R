# R front end of fitnoise_b200 over reticulate.
#
# Mirrors the interface of the reference's R wrapper (fitnoise/R/fitnoise.R:20-126: argument names,
# defaults, 1-based `coef`, and the fields of the returned lists) but is written for reticulate:
# matrices cross as numpy arrays instead of JSON strings, and the fitted Python object is kept as a
# reticulate reference in `$pyfit`.  UNTESTED in this repository's build image (no R available).

.fitnoise_env <- new.env()

.fitnoise_py <- function() {
    if (is.null(.fitnoise_env$module))
        .fitnoise_env$module <- reticulate::import("fitnoise_b200", convert = FALSE)
    .fitnoise_env$module
}

.as_float_matrix <- function(x) {
    if (is.null(x)) return(NULL)
    x <- as.matrix(x)
    storage.mode(x) <- "double"
    x[is.na(x)] <- NaN
    reticulate::np_array(x, dtype = "float64", order = "C")
}

.to_r <- function(x) reticulate::py_to_r(x)

.nan_to_na <- function(x) { x[is.nan(x)] <- NA; x }

# model: a Python expression naming the noise model, evaluated inside the fitnoise_b200 module,
# e.g. "Model_t()", "Model_t_patseq()", "Model_t_factors(1)".
fitnoise.fit <- function(data, design, model = "Model_t()", noise.design = NULL,
                         control.design = NULL, controls = NULL, weights = NULL, counts = NULL,
                         verbose = FALSE, force.param = NULL) {
    if (inherits(data, "EList")) {
        if (!is.null(data$weights)) { stopifnot(is.null(weights)); weights <- data$weights }
        if (!is.null(data$other$counts)) { stopifnot(is.null(counts)); counts <- data$other$counts }
        data <- data$E
    }
    if (is.null(noise.design)) noise.design <- design
    fn <- .fitnoise_py()
    context <- reticulate::dict()
    if (!is.null(weights)) context["weights"] <- .as_float_matrix(weights)
    if (!is.null(counts)) context["counts"] <- .as_float_matrix(counts)
    dataset <- fn$Dataset(.as_float_matrix(data), context)
    # Python's own eval with the package namespace as globals (reticulate::py_eval takes no namespace)
    builtins <- reticulate::import_builtins(convert = FALSE)
    prototype <- builtins$eval(model, reticulate::py_get_attr(fn, "__dict__"))
    fit <- prototype$fit(
        dataset,
        design = .as_float_matrix(design),
        noise_design = .as_float_matrix(noise.design),
        control_design = .as_float_matrix(control.design),
        controls = if (is.null(controls)) NULL else as.logical(controls),
        verbose = isTRUE(verbose),
        force_param = if (is.null(force.param)) NULL else as.numeric(force.param))
    coef <- .nan_to_na(.to_r(fit$coef))
    rownames(coef) <- rownames(data)
    colnames(coef) <- colnames(design)
    list(
        pyfit = fit,
        description = .to_r(fit$description),
        param = .to_r(fn$as_jsonic(fit$param)),
        df = .to_r(fit$df),
        deviance = .to_r(fit$deviance),
        score = .to_r(fit$score),
        coef = coef,
        noise.p.values = .nan_to_na(as.numeric(.to_r(fit$noise_p_values))),
        noise.combined.p.value = .to_r(fit$noise_combined_p_value),
        averages = .nan_to_na(as.numeric(.to_r(fit$averages))))
}

# coef: 1-based coefficient numbers (the Python side counts from zero), or contrasts: a matrix with
# one row per coefficient and one column per contrast.
fitnoise.test <- function(fit, coef = NULL, contrasts = NULL) {
    pycoef <- if (is.null(coef)) NULL else as.list(as.integer(coef) - 1L)
    pycontrasts <- if (is.null(contrasts)) NULL else .as_float_matrix(contrasts)
    tested <- fit$pyfit$test(coef = pycoef, contrasts = pycontrasts)
    fit$pyfit <- tested
    fit$description <- .to_r(tested$description)
    fit$contrasts <- .nan_to_na(.to_r(tested$contrasts))
    if (!is.null(contrasts) && !is.null(colnames(contrasts)))
        colnames(fit$contrasts) <- colnames(contrasts)
    fit$p.values <- .nan_to_na(as.numeric(.to_r(tested$p_values)))
    fit$q.values <- as.numeric(.to_r(tested$q_values))
    fit
}

# Weight matrix for limma (1 / variance of every observation under the fitted noise model).
fitnoise.weights <- function(fit) .to_r(fit$pyfit$get_weights())

# Table of a tested fit in the layout of the legacy test.fit (contrast columns, AveExpr, P.Value,
# adj.P.Val, noise.P.Value[, control]), ordered by p-value.
fitnoise.top.table <- function(fit, sort = TRUE) {
    table <- .to_r(fit$pyfit$top_table(sort = isTRUE(sort)))
    as.data.frame(table, check.names = FALSE)
}
