// inert stand-in
