"""Mirror of the reference's `model` package for the NDP path (nets, loss, registration)."""
