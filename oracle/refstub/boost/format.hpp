// TEST INFRASTRUCTURE: included by the reference's header bundle; nothing of it is used on the mapper path.
