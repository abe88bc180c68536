library(testthat)
library(plaid)
test_check("plaid")
