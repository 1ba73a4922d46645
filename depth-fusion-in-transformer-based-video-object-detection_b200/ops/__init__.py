"""Mirror of the reference's ``models/ops`` package (functions/ and modules/)."""
